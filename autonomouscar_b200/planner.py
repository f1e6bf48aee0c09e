"""Python face of the C ABI (include/prius_planner.h).

PyTorch is used for device memory, streams and torch.distributed only; every computation
is a call into libprius_planner_b200.so.  Method names follow the reference planner's
members they stand for (local_planner.cpp): set_grid <- costmapCallback/mapCostmapToCSpace,
collide <- checkCollision, plan <- localNavCallback/planPath.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .capi import (PP_GRID_CSPACE, PP_OK, PP_PHASE_COUNT, default_params, pp_params, pp_query, pp_result)
from .results import PlanBatch, PlanResult


class PlannerError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"pp status {code}: {text}")
        self.code = code


def _dptr(t):
    return C.c_void_p(t.data_ptr())


class Planner:
    """One planner instance on one GPU (pp_handle)."""

    def __init__(self, params=None, device=0):
        self._L = _lib.load()
        self.params = params if params is not None else default_params()
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        h = C.c_void_p()
        self._h = None
        self._check(self._L.pp_create(C.byref(self.params), self.device, C.byref(h)))
        self._h = h
        s = C.c_void_p()
        self._check(self._L.pp_stream(self._h, C.byref(s)))
        self.stream = torch.cuda.ExternalStream(s.value, device=self.torch_device)
        self._pending = {}     # tensors in use by calls still queued on self.stream (see _dev)

    def _dev(self, t):
        """Device pointer of a tensor handed to an ASYNCHRONOUS entry point.  The call runs on self.stream, which
        torch's caching allocator does not know about: a temporary freed right after the call could be handed out
        again (and overwritten on torch's stream) before the kernel has read it.  The planner therefore holds a
        reference until its stream is next synchronised (sync(), or here once a few thousand have piled up)."""
        if t.is_cuda:
            if len(self._pending) >= 4096:
                self.sync()
            self._pending[t.data_ptr()] = t
        return C.c_void_p(t.data_ptr())

    def _check(self, rc):
        if rc != PP_OK:
            raise PlannerError(rc, self._L.pp_last_error().decode())

    def close(self):
        if self._h is not None:
            self._L.pp_destroy(self._h)      # (synchronises the handle's stream)
            self._h = None
            self._pending.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- grid ingest -----------------------------------------------------------------
    def set_grid(self, data, layout=PP_GRID_CSPACE):
        W, H = self.params.costmap_width, self.params.costmap_height
        if isinstance(data, torch.Tensor) and data.is_cuda:
            assert data.dtype == torch.int8 and data.numel() == W * H and data.is_contiguous()
            torch.cuda.current_stream(self.torch_device).synchronize()
            self._check(self._L.pp_set_grid_dev(self._h, _dptr(data), W, H, self.params.costmap_res, layout))
        else:
            a = np.ascontiguousarray(np.asarray(data), dtype=np.int8)
            assert a.size == W * H, "grid size differs from costmap_width * costmap_height"
            self._check(self._L.pp_set_grid(self._h, a.ctypes.data_as(C.c_void_p), W, H,
                                            self.params.costmap_res, layout))

    def set_grid_bank_dev(self, grids_dev_ptr, n, layout=PP_GRID_CSPACE):
        """A bank of n grids of this planner's shape, already on the device, width*height bytes apart (e.g.
        LocalCostmap.batch_grids_dev()).  Queries name their grid with make_query(..., grid_index=k) and go
        through plan_batch(..., banked=True)."""
        self._check(self._L.pp_set_grid_bank_dev(self._h, int(n), C.c_void_p(grids_dev_ptr), self.params.costmap_width,
                                                 self.params.costmap_height, self.params.costmap_res, layout))

    def get_cspace(self):
        W, H = self.params.costmap_width, self.params.costmap_height
        out = np.zeros(W * H, dtype=np.int8)
        self._check(self._L.pp_get_cspace(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out.reshape(W, H)

    def get_visited(self):
        W, H = self.params.costmap_width, self.params.costmap_height
        out = np.zeros(W * H, dtype=np.uint8)
        self._check(self._L.pp_get_visited(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out.reshape(W, H)

    def set_collision_mode(self, mode):
        self._check(self._L.pp_set_collision_mode(self._h, mode))
        self.params.collision_mode = mode

    def set_broad_phase(self, enabled):
        self._check(self._L.pp_set_broad_phase(self._h, int(bool(enabled))))

    # ---- collision -------------------------------------------------------------------
    def collide(self, poses_host, out=None):
        """checkCollision for a batch; host buffers in, host flags out (copies inside)."""
        if isinstance(poses_host, torch.Tensor):
            assert not poses_host.is_cuda and poses_host.dtype == torch.float32 and poses_host.is_contiguous()
            n = poses_host.numel() // 3
            if out is None:
                out = torch.empty(n, dtype=torch.uint8, pin_memory=poses_host.is_pinned())
            self._check(self._L.pp_collide_batch(self._h, n, _dptr(poses_host), _dptr(out)))
            return out
        a = np.ascontiguousarray(poses_host, dtype=np.float32).reshape(-1, 3)
        if out is None:
            out = np.zeros(len(a), dtype=np.uint8)
        self._check(self._L.pp_collide_batch(self._h, len(a), a.ctypes.data_as(C.c_void_p),
                                             out.ctypes.data_as(C.c_void_p)))
        return out

    def collide_q16(self, poses_q16, quant, out=None):
        """pp_collide_batch_q16: int16 (qx, qy, qyaw) host poses + pp_pose_quant -> host flags (copies inside)."""
        if isinstance(poses_q16, torch.Tensor):
            assert not poses_q16.is_cuda and poses_q16.dtype == torch.int16 and poses_q16.is_contiguous()
            n = poses_q16.numel() // 3
            if out is None:
                out = torch.empty(n, dtype=torch.uint8, pin_memory=poses_q16.is_pinned())
            self._check(self._L.pp_collide_batch_q16(self._h, n, _dptr(poses_q16), C.byref(quant), _dptr(out)))
            return out
        a = np.ascontiguousarray(poses_q16, dtype=np.int16).reshape(-1, 3)
        if out is None:
            out = np.zeros(len(a), dtype=np.uint8)
        self._check(self._L.pp_collide_batch_q16(self._h, len(a), a.ctypes.data_as(C.c_void_p), C.byref(quant),
                                                 out.ctypes.data_as(C.c_void_p)))
        return out

    def collide_dev(self, poses, out=None):
        """Device tensors; asynchronous on self.stream."""
        assert poses.is_cuda and poses.dtype == torch.float32 and poses.is_contiguous()
        n = poses.numel() // 3
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=poses.device)
        self._check(self._L.pp_collide_batch_dev(self._h, n, self._dev(poses), self._dev(out)))
        return out

    def collide_bits_dev(self, poses, out=None):
        """collide_dev with one bit per pose as the result (int32 words, bit b of word k = pose 32k + b): what
        pack_flags_dev(collide_dev(poses)) returns, written by the batch kernel itself in ORIENTED mode."""
        assert poses.is_cuda and poses.dtype == torch.float32 and poses.is_contiguous()
        n = poses.numel() // 3
        if out is None:
            out = torch.empty((n + 31) // 32, dtype=torch.int32, device=poses.device)
        self._check(self._L.pp_collide_batch_bits_dev(self._h, n, self._dev(poses), self._dev(out)))
        return out

    def pack_flags_dev(self, flags, out=None):
        """1 B/pose flags -> 1 bit/pose (int32 words, bit b of word k = flag 32k + b); asynchronous on self.stream."""
        n = flags.numel()
        if out is None:
            out = torch.empty((n + 31) // 32, dtype=torch.int32, device=flags.device)
        self._check(self._L.pp_pack_flags_dev(self._h, n, self._dev(flags), self._dev(out)))
        return out

    def unpack_flags_dev(self, bits, n, out=None):
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=bits.device)
        self._check(self._L.pp_unpack_flags_dev(self._h, n, self._dev(bits), self._dev(out)))
        return out

    def collide_edges_dev(self, edges, out=None):
        assert edges.is_cuda and edges.dtype == torch.float32 and edges.is_contiguous()
        n = edges.numel() // 5
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=edges.device)
        self._check(self._L.pp_collide_edges_dev(self._h, n, self._dev(edges), self._dev(out)))
        return out

    def footprint_points(self, x, y, yaw, capacity=4096):
        out = np.zeros((capacity, 2), dtype=np.float64)
        n = C.c_int32()
        self._check(self._L.pp_footprint_points(self._h, x, y, yaw, out.ctypes.data_as(C.c_void_p), capacity,
                                                C.byref(n)))
        return out[: n.value]

    # ---- sampler / nearest neighbour ---------------------------------------------------
    def sample_poses_dev(self, seed, stream, first, n, x_lo, x_hi, y_lo, y_hi, out=None):
        if out is None:
            out = torch.empty((n, 3), dtype=torch.float32, device=self.torch_device)
        self._check(self._L.pp_sample_poses_dev(self._h, seed, stream, first, n, x_lo, x_hi, y_lo, y_hi,
                                                self._dev(out)))
        return out

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self._check(self._L.pp_philox4x32_10(self._h, c, k, o))
        return [int(v) for v in o]

    def nearest_dev(self, node_xy, query_xy):
        assert node_xy.is_cuda and query_xy.is_cuda and node_xy.dtype == torch.float32
        nq = query_xy.numel() // 2
        idx = torch.empty(nq, dtype=torch.int32, device=query_xy.device)
        d2 = torch.empty(nq, dtype=torch.float32, device=query_xy.device)
        self._check(self._L.pp_nearest_dev(self._h, node_xy.numel() // 2, self._dev(node_xy), nq, self._dev(query_xy),
                                           self._dev(idx), self._dev(d2)))
        return idx, d2

    def find_exact_dev(self, keys, queries):
        assert keys.is_cuda and queries.is_cuda and keys.dtype == torch.float64
        nq = queries.numel() // 6
        idx = torch.empty(nq, dtype=torch.int32, device=queries.device)
        self._check(self._L.pp_find_exact_dev(self._h, keys.numel() // 6, self._dev(keys), nq, self._dev(queries),
                                              self._dev(idx)))
        return idx

    # ---- planning ----------------------------------------------------------------------
    def plan(self, query, **caps):
        res = PlanResult(**caps)
        res.rc = self._L.pp_plan(self._h, C.byref(query), C.byref(res.raw))
        if res.rc not in (PP_OK, 4):
            self._check(res.rc)
        return res

    def plan_batch(self, queries, want_nodes=False, banked=False, **caps):
        n = len(queries)
        qs = (pp_query * n)(*queries)
        results = [PlanResult(want_nodes=want_nodes, **caps) for _ in range(n)]
        raw = (pp_result * n)(*[r.raw for r in results])
        rc = (self._L.pp_plan_batch_banked if banked else self._L.pp_plan_batch)(self._h, n, qs, raw)
        if rc not in (PP_OK, 4):
            self._check(rc)
        for r, rr in zip(results, raw):
            r.raw = rr
        return results

    def plan_batch_into(self, queries_array, batch, rrt=False, banked=False):
        """pp_plan_batch / pp_rrt_plan_batch / pp_plan_batch_banked into pre-allocated PlanBatch buffers.
        `queries_array` is a ctypes (pp_query * n) array."""
        fn = self._L.pp_rrt_plan_batch if rrt else (self._L.pp_plan_batch_banked if banked else self._L.pp_plan_batch)
        rc = fn(self._h, batch.n, queries_array, batch.raw)
        if rc not in (PP_OK, 4):
            self._check(rc)
        return batch

    def plan_batch_dev(self, queries_array, n, cap_path=128, out=None):
        """pp_plan_batch_dev: n queries (ctypes pp_query array, host) -> fixed-stride records that STAY on the device
        (uint8 tensor [n, pp_plan_record_stride(cap_path)]).  Asynchronous on self.stream; results.PlanRecords views
        a host copy."""
        stride = int(self._L.pp_plan_record_stride(int(cap_path)))
        if out is None:
            out = torch.empty((n, stride), dtype=torch.uint8, device=self.torch_device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.numel() >= n * stride
        self._check(self._L.pp_plan_batch_dev(self._h, int(n), queries_array, int(cap_path), self._dev(out)))
        return out

    def rrt_plan(self, query, **caps):
        res = PlanResult(**caps)
        res.rc = self._L.pp_rrt_plan(self._h, C.byref(query), C.byref(res.raw))
        if res.rc not in (PP_OK, 4):
            self._check(res.rc)
        return res

    def rrt_plan_batch(self, queries, want_nodes=False, **caps):
        n = len(queries)
        qs = (pp_query * n)(*queries)
        results = [PlanResult(want_nodes=want_nodes, **caps) for _ in range(n)]
        raw = (pp_result * n)(*[r.raw for r in results])
        rc = self._L.pp_rrt_plan_batch(self._h, n, qs, raw)
        if rc not in (PP_OK, 4):
            self._check(rc)
        for r, rr in zip(results, raw):
            r.raw = rr
        return results

    # ---- plumbing ------------------------------------------------------------------------
    def sync(self):
        self._check(self._L.pp_sync(self._h))
        self._pending.clear()

    def launch_count(self):
        v = C.c_uint64()
        self._check(self._L.pp_launch_count(self._h, C.byref(v)))
        return v.value

    def last_phase_ms(self):
        a = (C.c_float * PP_PHASE_COUNT)()
        self._check(self._L.pp_last_phase_ms(self._h, a, PP_PHASE_COUNT))
        return [float(v) for v in a]

    def bench_l2_gather(self, buffer_bytes=16 << 20):
        v = C.c_double()
        self._check(self._L.pp_bench_l2_gather(self._h, buffer_bytes, C.byref(v)))
        return v.value

    def last_narrow_count(self):
        v = C.c_int64()
        self._check(self._L.pp_last_narrow_count(self._h, C.byref(v)))
        return v.value


class PlannerGroup:
    """Single-process multi-GPU front end (pp_group_*): one handle per device, pose batches
    and plan queries sharded contiguously, results gathered with NCCL."""

    def __init__(self, params=None, devices=None):
        self._L = _lib.load()
        self.params = params if params is not None else default_params()
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        g = C.c_void_p()
        self._g = None
        rc = self._L.pp_group_create(C.byref(self.params), len(self.devices), arr, C.byref(g))
        if rc != PP_OK:
            raise PlannerError(rc, self._L.pp_last_error().decode())
        self._g = g

    def _check(self, rc):
        if rc != PP_OK:
            raise PlannerError(rc, self._L.pp_last_error().decode())

    def close(self):
        if self._g is not None:
            self._L.pp_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        n = C.c_int()
        self._check(self._L.pp_group_size(self._g, C.byref(n)))
        return n.value

    def set_grid(self, data, layout=PP_GRID_CSPACE):
        a = np.ascontiguousarray(np.asarray(data), dtype=np.int8)
        self._check(self._L.pp_group_set_grid(self._g, a.ctypes.data_as(C.c_void_p), self.params.costmap_width,
                                              self.params.costmap_height, self.params.costmap_res, layout))

    def collide(self, poses_host):
        a = np.ascontiguousarray(poses_host, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(a), dtype=np.uint8)
        self._check(self._L.pp_group_collide_batch(self._g, len(a), a.ctypes.data_as(C.c_void_p),
                                                   out.ctypes.data_as(C.c_void_p)))
        return out

    def records_dev(self, device_index, n, cap_path):
        """Device `device_index`'s copy of the records gathered by the last plan_batch, as a zero-copy uint8 tensor
        [n, stride] (valid until the next call on the group)."""
        ptr = C.c_void_p()
        self._check(self._L.pp_group_records_dev(self._g, device_index, C.byref(ptr)))
        stride = int(self._L.pp_plan_record_stride(int(cap_path)))

        class _View:
            __cuda_array_interface__ = {"shape": (n * stride,), "typestr": "|u1", "data": (ptr.value, False), "version": 2}
        dev = torch.device("cuda", self.devices[device_index])
        return torch.as_tensor(_View(), device=dev).reshape(n, stride)

    def plan_batch(self, queries, want_nodes=False, **caps):
        n = len(queries)
        qs = (pp_query * n)(*queries)
        results = [PlanResult(want_nodes=want_nodes, **caps) for _ in range(n)]
        if not want_nodes:       # paths + profiles only: the device-resident records + NCCL gather path
            for r in results:
                r.raw.expansion_ids = None
                r.raw.cap_expansions = 0
        raw = (pp_result * n)(*[r.raw for r in results])
        rc = self._L.pp_group_plan_batch(self._g, n, qs, raw)
        if rc not in (PP_OK, 4):
            self._check(rc)
        for r, rr in zip(results, raw):
            r.raw = rr
        return results
