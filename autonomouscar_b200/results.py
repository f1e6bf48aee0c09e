"""Host-side buffers behind a pp_result (numpy arrays the C side fills in)."""
import ctypes as C

import numpy as np

from .capi import pp_result


class PlanResult:
    """Owns the output arrays of one plan and exposes them trimmed to the filled length."""

    def __init__(self, cap_nodes=16384, cap_path=4096, cap_expansions=16384, want_nodes=True):
        self.raw = pp_result()
        self.raw.cap_nodes = cap_nodes
        self.raw.cap_path = cap_path
        self.raw.cap_expansions = cap_expansions
        self._node_state = np.zeros((cap_nodes, 8), dtype=np.float64) if want_nodes else None
        self._node_parent = np.full(cap_nodes, -2, dtype=np.int32) if want_nodes else None
        self._node_closed_seq = np.full(cap_nodes, -2, dtype=np.int32) if want_nodes else None
        self._expansion_ids = np.full(cap_expansions, -2, dtype=np.int32)
        self._path_xy = np.zeros((cap_path, 2), dtype=np.float64)
        self._path_node_ids = np.full(cap_path, -2, dtype=np.int32)
        self._yaw_rates = np.zeros(cap_path, dtype=np.float64)
        self._durations = np.zeros(cap_path, dtype=np.float64)
        self._speeds = np.zeros(cap_path, dtype=np.float64)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        if want_nodes:
            self.raw.node_state = self._node_state.ctypes.data_as(dp)
            self.raw.node_parent = self._node_parent.ctypes.data_as(ip)
            self.raw.node_closed_seq = self._node_closed_seq.ctypes.data_as(ip)
        self.raw.expansion_ids = self._expansion_ids.ctypes.data_as(ip)
        self.raw.path_xy = self._path_xy.ctypes.data_as(dp)
        self.raw.path_node_ids = self._path_node_ids.ctypes.data_as(ip)
        self.raw.yaw_rates = self._yaw_rates.ctypes.data_as(dp)
        self.raw.durations = self._durations.ctypes.data_as(dp)
        self.raw.speeds = self._speeds.ctypes.data_as(dp)

    # scalars -------------------------------------------------------------------------
    def __getattr__(self, name):
        if name in ("status", "n_expansions", "n_nodes", "n_open", "n_closed", "n_path", "n_profile",
                    "n_candidates", "n_collisions", "tree_digest"):
            return getattr(self.raw, name)
        raise AttributeError(name)

    # arrays --------------------------------------------------------------------------
    @property
    def node_state(self):
        return self._node_state[: min(self.raw.n_nodes, self.raw.cap_nodes)]

    @property
    def node_parent(self):
        return self._node_parent[: min(self.raw.n_nodes, self.raw.cap_nodes)]

    @property
    def node_closed_seq(self):
        return self._node_closed_seq[: min(self.raw.n_nodes, self.raw.cap_nodes)]

    @property
    def expansion_ids(self):
        return self._expansion_ids[: min(self.raw.n_expansions, self.raw.cap_expansions)]

    @property
    def path_xy(self):
        return self._path_xy[: min(self.raw.n_path, self.raw.cap_path)]

    @property
    def path_node_ids(self):
        return self._path_node_ids[: min(self.raw.n_path, self.raw.cap_path)]

    @property
    def yaw_rates(self):
        return self._yaw_rates[: min(self.raw.n_profile, self.raw.cap_path)]

    @property
    def durations(self):
        return self._durations[: min(self.raw.n_profile, self.raw.cap_path)]

    @property
    def speeds(self):
        return self._speeds[: min(self.raw.n_profile, self.raw.cap_path)]

    def summary(self):
        return {k: int(getattr(self.raw, k)) for k in (
            "status", "n_expansions", "n_nodes", "n_open", "n_closed", "n_path", "n_profile",
            "n_candidates", "n_collisions", "tree_digest")}


class PlanBatch:
    """Output buffers of a whole batch of plans in a few contiguous arrays (one row per
    query), allocated once and reusable across calls -- the fast path of pp_plan_batch /
    pp_rrt_plan_batch when only paths, profiles and summaries are wanted."""

    def __init__(self, n, cap_path=256):
        self.n, self.cap_path = n, cap_path
        self.path_xy = np.zeros((n, cap_path, 2), dtype=np.float64)
        self.path_node_ids = np.full((n, cap_path), -2, dtype=np.int32)
        self.yaw_rates = np.zeros((n, cap_path), dtype=np.float64)
        self.durations = np.zeros((n, cap_path), dtype=np.float64)
        self.speeds = np.zeros((n, cap_path), dtype=np.float64)
        self.raw = (pp_result * n)()
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        for k in range(n):
            r = self.raw[k]
            r.cap_nodes, r.cap_path, r.cap_expansions = 0, cap_path, 0
            r.path_xy = self.path_xy[k].ctypes.data_as(dp)
            r.path_node_ids = self.path_node_ids[k].ctypes.data_as(ip)
            r.yaw_rates = self.yaw_rates[k].ctypes.data_as(dp)
            r.durations = self.durations[k].ctypes.data_as(dp)
            r.speeds = self.speeds[k].ctypes.data_as(dp)

    def _view(self):
        """The pp_result array as a numpy structured array (no copy)."""
        if getattr(self, "_np_view", None) is None:
            names, formats, offsets = [], [], []
            for name, ctype in pp_result._fields_:
                names.append(name)
                formats.append(np.dtype(ctype) if not hasattr(ctype, "contents") else np.dtype(np.uintp))
                offsets.append(getattr(pp_result, name).offset)
            dt = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": C.sizeof(pp_result)})
            self._np_view = np.frombuffer(self.raw, dtype=dt, count=self.n)
        return self._np_view

    def field(self, name):
        return np.array(self._view()[name])

    def records(self):
        """[n, 5] int64: status, n_expansions, n_nodes, n_path, digest (low 63 bits)."""
        v = self._view()
        out = np.empty((self.n, 5), dtype=np.int64)
        out[:, 0], out[:, 1], out[:, 2], out[:, 3] = v["status"], v["n_expansions"], v["n_nodes"], v["n_path"]
        out[:, 4] = (v["tree_digest"] & np.uint64(0x7FFFFFFFFFFFFFFF)).astype(np.int64)
        return out


RECORD_HEADER = np.dtype([("status", "<i4"), ("n_expansions", "<i4"), ("n_nodes", "<i4"), ("n_open", "<i4"),
                          ("n_closed", "<i4"), ("n_path", "<i4"), ("n_profile", "<i4"), ("n_candidates", "<i4"),
                          ("n_collisions", "<i4"), ("truncated", "<i4"), ("cap_path", "<i4"), ("reserved", "<i4"),
                          ("tree_digest", "<u8"), ("reserved2", "<u8")])


def record_stride(cap_path):
    """pp_plan_record_stride: header (64 B) + 44 B per path slot, rounded up to 16 B."""
    return (RECORD_HEADER.itemsize + 44 * int(cap_path) + 15) & ~15


def record_dtype(cap_path):
    """numpy view of one pp_plan_batch_dev record (include/prius_planner.h: pp_plan_record_header | path_xy |
    yaw_rates | durations | speeds | path_node_ids)."""
    c = int(cap_path)
    o = RECORD_HEADER.itemsize
    return np.dtype({"names": ["hdr", "path_xy", "yaw_rates", "durations", "speeds", "path_node_ids"],
                     "formats": [RECORD_HEADER, ("<f8", (c, 2)), ("<f8", (c,)), ("<f8", (c,)), ("<f8", (c,)), ("<i4", (c,))],
                     "offsets": [0, o, o + 16 * c, o + 24 * c, o + 32 * c, o + 40 * c],
                     "itemsize": record_stride(c)})


class PlanRecords:
    """Host view of a block of fixed-stride plan records (a uint8 array [n, stride] copied from the device, e.g.
    the gathered block of sharding.plan_batch_gather)."""

    def __init__(self, raw, cap_path):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        self.cap_path = int(cap_path)
        self.rec = raw.reshape(-1).view(record_dtype(cap_path))
        self.n = len(self.rec)

    def field(self, name):
        return self.rec["hdr"][name]

    def path_xy(self, k):
        return self.rec["path_xy"][k][: min(int(self.rec["hdr"]["n_path"][k]), self.cap_path)]

    def path_node_ids(self, k):
        return self.rec["path_node_ids"][k][: min(int(self.rec["hdr"]["n_path"][k]), self.cap_path)]

    def profile(self, k):
        m = min(int(self.rec["hdr"]["n_profile"][k]), self.cap_path)
        return self.rec["yaw_rates"][k][:m], self.rec["durations"][k][:m], self.rec["speeds"][k][:m]

    def summary(self, k):
        h = self.rec["hdr"][k]
        return {f: int(h[f]) for f in ("status", "n_expansions", "n_nodes", "n_open", "n_closed", "n_path", "n_profile",
                                        "n_candidates", "n_collisions", "tree_digest")}
