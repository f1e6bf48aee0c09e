"""The oracle restatements under AddressSanitizer + UndefinedBehaviorSanitizer (tests/cpp/test_oracle_sanitize.cpp):
where the reference has undefined behaviour (unchecked grid indices, top() of an empty queue, out-of-range
double -> int conversions, NaN inputs) the oracle -- and with it the contract of the CUDA path -- is defined."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_clean_under_asan_ubsan(tmp_path):
    exe = str(tmp_path / "oracle_sanitize")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_oracle_sanitize.cpp"), "-o", exe])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr[-3000:]
    assert "found" in out.stdout and "runtime error" not in out.stderr
