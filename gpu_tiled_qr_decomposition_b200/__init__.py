"""gpu_tiled_qr_decomposition_b200 -- ctypes binding of libqrb200.so.

The product is the C-ABI shared library (include/qrb200.h, include/gpucalc_compat.h) built
from csrc/ for sm_100a; this module only loads it for the Python test/bench harness.  There
is NO CPU fallback: if the library is missing, or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QRB200_LIB") or os.path.join(HERE, "libqrb200.so")   # QRB200_LIB: A/B runs of kernel variants

TILE = 32
QRS, SAPP, QRD, DAPP = 0, 1, 2, 3
TASK_AVAIL, TASK_NONE, TASK_DONE = 0, 1, 2

# every symbol include/qrb200.h and include/gpucalc_compat.h declare
EXPORTS = (
    "b200qr_version", "b200qr_strerror", "b200qr_last_cuda_error", "b200qr_device_sm_count", "b200qr_copy",
    "b200qr_measure_fp32_peak",
    "b200qr_plan_create", "b200qr_plan_destroy", "b200qr_plan_grid", "b200qr_plan_variant", "b200qr_plan_debug_counters", "b200qr_plan_tile_tasks",
    "b200qr_plan_macro_tasks", "b200qr_factor_device", "b200qr_plan_tau_device",
    "b200qr_tau_to_reference", "b200qr_factor_host", "b200qr_plan_read_trace", "b200qr_tile_task",
    "b200qr_sched_init", "b200qr_sched_next", "b200qr_sched_done", "b200qr_sched_free",
    "b200qr_sched_progress", "b200qr_sched_final_progress", "b200qr_sched_ready_count", "b200qr_sched_dep_count",
    "b200qr_sched_predecessors", "b200qr_sched_device_order", "b200qr_sched_sweep_order", "b200qr_sched_rank_tickets", "b200qr_sched_sweep_order_skew",
    "b200qr_sched_rank_tickets_skew", "b200qr_sched_sweep_order_lag",
    "b200qr_dist_create", "b200qr_dist_destroy", "b200qr_dist_handle_bytes", "b200qr_dist_export",
    "b200qr_dist_connect", "b200qr_dist_matrix", "b200qr_dist_tau_device", "b200qr_dist_block",
    "b200qr_dist_owner", "b200qr_dist_grid", "b200qr_dist_variant", "b200qr_dist_read_trace", "b200qr_dist_tile_tasks", "b200qr_dist_reset", "b200qr_dist_factor", "b200qr_dist_reset_host", "b200qr_dist_factor_host", "b200qr_plan_stall_info", "b200qr_plan_stall_notes", "b200qr_dist_stall_info",
    "b200qr_dist_connect_local", "b200qr_dist_factor_group",
    "b200qr_batched_workspace_floats", "b200qr_batched_factor", "b200qr_batched_last_cuda_error",
    "b200qr_tsqr_create", "b200qr_tsqr_create_streaming", "b200qr_tsqr_create_tiles", "b200qr_tsqr_leaves", "b200qr_tsqr_stream_block_rows", "b200qr_tsqr_create_rstack", "b200qr_tsqr_destroy", "b200qr_tsqr_levels", "b200qr_tsqr_factor",
    "b200qr_tsqr_dist_create", "b200qr_tsqr_dist_destroy", "b200qr_tsqr_dist_handle_bytes", "b200qr_tsqr_dist_export",
    "b200qr_tsqr_dist_connect", "b200qr_tsqr_dist_connect_local", "b200qr_tsqr_dist_factor", "b200qr_tsqr_dist_factor_begin",
    "b200qr_tsqr_dist_factor_end", "b200qr_tsqr_dist_local_r",
    "b200qr_apply_q", "b200qr_form_q", "b200qr_trsm_upper", "b200qr_solve_ls",
    "b200qr_any_create", "b200qr_any_destroy", "b200qr_any_padded_rows", "b200qr_any_padded_cols", "b200qr_any_factor",
    "b200qr_any_get_r", "b200qr_any_apply_q", "b200qr_any_solve_ls", "b200qr_any_factored", "b200qr_any_tau",
    "b200qr_frontend_last_cuda_error",
    "b200qr_tiles_create", "b200qr_tiles_destroy", "b200qr_tiles_tile", "b200qr_tiles_factor", "b200qr_tiles_tau",
    "b200qr_tiles_apply_q", "b200qr_tiles_last_cuda_error",
    "b200qr_release_cache", "cudaQRTask", "cudaQRFull", "testDAPP", "doCUDADAPP",
)


class QRError(RuntimeError):
    pass


class Task(C.Structure):
    _fields_ = [("type", C.c_int), ("l", C.c_int), ("m", C.c_int), ("k", C.c_int)]


class Opts(C.Structure):
    _fields_ = [("grid_ctas", C.c_int), ("trace", C.c_int), ("share_device", C.c_int), ("engine", C.c_int),
                ("reserved", C.c_int * 4)]


PROBE_LIB_PATH = os.path.join(HERE, "libqrb200_probe.so")     # diagnostics (tools/qrb200_probe.h), not the product ABI


def build(force=False):
    """Compile libqrb200.so (and the diagnostics library) in-tree with nvcc for sm_100a (no GPU needed)."""
    srcs = []
    for d, _, fs in os.walk(os.path.join(HERE, "csrc")):
        srcs += [os.path.join(d, f) for f in fs]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "qrb200.h"))
    srcs.append(os.path.join(os.path.dirname(HERE), "tools", "qrb200_probe.h"))
    srcs.append(os.path.join(HERE, "Makefile"))
    here_lib = os.path.join(HERE, "libqrb200.so")
    stale = force or any(not os.path.exists(l) or any(os.path.getmtime(s) > os.path.getmtime(l) for s in srcs)
                         for l in (here_lib, PROBE_LIB_PATH))
    if stale:
        r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
        if r.returncode != 0:
            raise QRError("building libqrb200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_probe = None


def probe_lib():
    """The diagnostics library (panel-latency and tcgen05 probes)."""
    global _probe
    if _probe is None:
        if not os.path.exists(PROBE_LIB_PATH):
            raise QRError("libqrb200_probe.so is not built")
        P = C.CDLL(PROBE_LIB_PATH)
        vp = C.c_void_p
        P.b200qr_probe_panel.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), vp]
        P.b200qr_probe_dram_pattern.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        P.b200qr_probe_sweep.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        P.b200qr_probe_umma.argtypes = [vp, vp, vp, vp, vp]
        P.b200qr_probe_umma_mode.argtypes = [vp, vp, vp, vp, vp, C.c_int]
        P.b200qr_probe_umma_rate.argtypes = [C.POINTER(C.c_longlong), C.c_int]
        P.b200qr_probe_umma_rate2.argtypes = [C.POINTER(C.c_longlong), C.c_int]
        _probe = P
    return _probe


_lib = None


def lib():
    """The loaded library (raises QRError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QRError("libqrb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.b200qr_strerror.restype = C.c_char_p
        L.b200qr_strerror.argtypes = [C.c_int]
        L.b200qr_copy.argtypes = [vp, vp, C.c_size_t]
        L.b200qr_measure_fp32_peak.argtypes = [C.POINTER(C.c_double), C.c_int]
        L.b200qr_plan_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(Opts)]
        L.b200qr_plan_destroy.argtypes = [vp]
        L.b200qr_plan_grid.argtypes = [vp]
        L.b200qr_plan_variant.argtypes = [vp]
        L.b200qr_plan_debug_counters.argtypes = [vp, vp, C.c_int]
        L.b200qr_plan_tile_tasks.argtypes = [vp]
        L.b200qr_plan_tile_tasks.restype = C.c_long
        L.b200qr_plan_macro_tasks.argtypes = [vp]
        L.b200qr_plan_macro_tasks.restype = C.c_long
        L.b200qr_factor_device.argtypes = [vp, vp, C.c_int, vp]
        L.b200qr_plan_tau_device.argtypes = [vp]
        L.b200qr_plan_tau_device.restype = vp
        L.b200qr_tau_to_reference.argtypes = [vp, vp, C.c_int, vp]
        L.b200qr_factor_host.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, fp]
        L.b200qr_plan_read_trace.argtypes = [vp, vp, C.c_long]
        L.b200qr_tile_task.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        L.b200qr_sched_init.argtypes = [C.c_int, C.c_int]
        L.b200qr_sched_init.restype = vp
        L.b200qr_sched_next.argtypes = [vp, C.POINTER(Task)]
        L.b200qr_sched_done.argtypes = [vp, C.POINTER(Task)]
        L.b200qr_sched_free.argtypes = [vp]
        L.b200qr_sched_progress.argtypes = [vp, C.c_int, C.c_int]
        L.b200qr_sched_ready_count.argtypes = [vp]
        L.b200qr_sched_dep_count.argtypes = [C.c_int, C.c_int]
        L.b200qr_sched_predecessors.argtypes = [C.POINTER(Task), C.POINTER(Task)]
        L.b200qr_sched_device_order.argtypes = [C.c_int, C.c_int, vp, C.c_long]
        L.b200qr_sched_device_order.restype = C.c_long
        L.b200qr_sched_sweep_order.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_long]
        L.b200qr_sched_sweep_order.restype = C.c_long
        L.b200qr_sched_sweep_order_lag.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_long]
        L.b200qr_sched_sweep_order_lag.restype = C.c_long
        L.b200qr_sched_rank_tickets.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp, C.c_long,
                                                C.POINTER(C.c_long)]
        L.b200qr_sched_rank_tickets.restype = C.c_long
        L.b200qr_dist_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Opts)]
        L.b200qr_dist_destroy.argtypes = [vp]
        L.b200qr_dist_export.argtypes = [vp, vp]
        L.b200qr_dist_connect.argtypes = [vp, vp]
        L.b200qr_dist_matrix.argtypes = [vp]
        L.b200qr_dist_matrix.restype = vp
        L.b200qr_dist_tau_device.argtypes = [vp]
        L.b200qr_dist_tau_device.restype = vp
        L.b200qr_dist_block.argtypes = [vp]
        L.b200qr_dist_owner.argtypes = [vp, C.c_int]
        L.b200qr_dist_grid.argtypes = [vp]
        L.b200qr_dist_reset.argtypes = [vp, vp]
        L.b200qr_dist_read_trace.argtypes = [vp, vp, C.c_long]
        L.b200qr_dist_tile_tasks.argtypes = [vp]
        L.b200qr_dist_tile_tasks.restype = C.c_long
        L.b200qr_dist_factor.argtypes = [vp, vp]
        L.b200qr_dist_reset_host.argtypes = [vp]
        L.b200qr_plan_stall_info.argtypes = [vp, vp]
        L.b200qr_plan_stall_notes.argtypes = [vp, vp, C.c_int]
        L.b200qr_dist_stall_info.argtypes = [vp, vp]
        L.b200qr_dist_factor_host.argtypes = [vp, vp, C.c_longlong, C.c_int, C.POINTER(C.c_float)]
        L.b200qr_dist_connect_local.argtypes = [C.POINTER(vp), C.c_int]
        L.b200qr_dist_factor_group.argtypes = [C.POINTER(vp), C.c_int, vp]
        L.b200qr_batched_workspace_floats.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        L.b200qr_batched_factor.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, vp, vp, vp]
        L.b200qr_tsqr_create.argtypes = [C.POINTER(vp), C.c_longlong, C.c_int, C.c_int, C.c_int]
        L.b200qr_tsqr_create_streaming.argtypes = [C.POINTER(vp), C.c_longlong, C.c_int, C.c_int]
        L.b200qr_tsqr_create_tiles.argtypes = [C.POINTER(vp), C.c_longlong, C.c_int, C.c_int]
        L.b200qr_tsqr_leaves.argtypes = [vp]
        L.b200qr_tsqr_stream_block_rows.argtypes = [vp]
        L.b200qr_tsqr_create_rstack.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
        L.b200qr_tsqr_destroy.argtypes = [vp]
        L.b200qr_tsqr_levels.argtypes = [vp]
        L.b200qr_tsqr_factor.argtypes = [vp, vp, C.c_longlong, vp, vp]
        L.b200qr_tsqr_dist_create.argtypes = [C.POINTER(vp), C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]
        L.b200qr_tsqr_dist_destroy.argtypes = [vp]
        L.b200qr_tsqr_dist_export.argtypes = [vp, vp]
        L.b200qr_tsqr_dist_connect.argtypes = [vp, vp]
        L.b200qr_tsqr_dist_connect_local.argtypes = [C.POINTER(vp), C.c_int]
        L.b200qr_tsqr_dist_factor.argtypes = [vp, vp, C.c_longlong, vp, vp]
        L.b200qr_tsqr_dist_factor_begin.argtypes = [vp, vp, C.c_longlong, vp]
        L.b200qr_tsqr_dist_factor_end.argtypes = [vp, vp, vp]
        L.b200qr_tsqr_dist_local_r.argtypes = [vp]
        L.b200qr_tsqr_dist_local_r.restype = vp
        L.b200qr_release_cache.restype = None
        L.b200qr_apply_q.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp]
        L.b200qr_form_q.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]
        L.b200qr_trsm_upper.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp]
        L.b200qr_solve_ls.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp]
        L.b200qr_any_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(Opts)]
        L.b200qr_any_destroy.argtypes = [vp]
        L.b200qr_any_padded_rows.argtypes = [vp]
        L.b200qr_any_padded_cols.argtypes = [vp]
        L.b200qr_any_factor.argtypes = [vp, vp, C.c_int, vp]
        L.b200qr_any_get_r.argtypes = [vp, vp, C.c_int, vp]
        L.b200qr_any_apply_q.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
        L.b200qr_any_solve_ls.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]
        L.b200qr_any_factored.argtypes = [vp]
        L.b200qr_any_factored.restype = vp
        L.b200qr_any_tau.argtypes = [vp]
        L.b200qr_any_tau.restype = vp
        L.b200qr_tiles_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
        L.b200qr_tiles_destroy.argtypes = [vp]
        L.b200qr_tiles_tile.argtypes = [vp]
        L.b200qr_tiles_factor.argtypes = [vp, vp, C.c_int, vp]
        L.b200qr_tiles_tau.argtypes = [vp]
        L.b200qr_tiles_tau.restype = vp
        L.b200qr_tiles_apply_q.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]
        L.cudaQRTask.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.cudaQRTask.restype = None
        L.cudaQRFull.argtypes = [vp, C.c_int, C.c_int]
        L.cudaQRFull.restype = None
        L.testDAPP.argtypes = [vp, C.c_int, C.c_int]
        L.testDAPP.restype = None
        L.doCUDADAPP.argtypes = [vp]
        L.doCUDADAPP.restype = None
        _lib = L
    return _lib


def check(rc, what="libqrb200"):
    if rc != 0:
        L = lib()
        raise QRError("%s failed: %s (cuda error %d)" % (what, L.b200qr_strerror(rc).decode(),
                                                        L.b200qr_last_cuda_error()))


class Plan:
    """Workspace + schedule for one (m, n) shape (b200qr_plan_*)."""

    def __init__(self, m, n, grid_ctas=0, trace=False, engine=0):
        """engine: 0 = automatic, 1 = FP32 SIMT sweeps, 2 = tensor-core sweeps (b200qr_opts.engine)."""
        self.m, self.n = m, n
        self._h = C.c_void_p()
        o = Opts()
        o.grid_ctas, o.trace, o.engine = int(grid_ctas), int(trace), int(engine)
        check(lib().b200qr_plan_create(C.byref(self._h), m, n, C.byref(o)), "b200qr_plan_create")

    def stall_info(self):
        """After a stall (launch failure): (valid, block, warp, tag, arg) of the first wait that gave up."""
        out = (C.c_uint * 5)()
        check(lib().b200qr_plan_stall_info(self._h, out), "b200qr_plan_stall_info")
        return tuple(out)

    def stall_notes(self):
        """{(block, warp): (tag, arg)} of every warp that gave up waiting before the kernel trapped."""
        cap = 16 * 256
        out = (C.c_uint * (2 * cap))()
        n = lib().b200qr_plan_stall_notes(self._h, out, cap)
        return {(s // 16, s % 16): (out[2 * s], out[2 * s + 1] & 0x7fffffff) for s in range(max(n, 0)) if out[2 * s + 1]}

    @property
    def grid(self):
        return lib().b200qr_plan_grid(self._h)

    @property
    def variant(self):
        return lib().b200qr_plan_variant(self._h)

    @property
    def tile_tasks(self):
        return lib().b200qr_plan_tile_tasks(self._h)

    @property
    def macro_tasks(self):
        return lib().b200qr_plan_macro_tasks(self._h)

    def factor(self, dptr, ldm, stream=0):
        """Factor the device matrix at address `dptr` in place (asynchronous on `stream`)."""
        check(lib().b200qr_factor_device(self._h, C.c_void_p(dptr), ldm, C.c_void_p(stream)),
              "b200qr_factor_device")

    def tau_to_reference(self, dptr, ldtau, stream=0):
        check(lib().b200qr_tau_to_reference(self._h, C.c_void_p(dptr), ldtau, C.c_void_p(stream)),
              "b200qr_tau_to_reference")

    @property
    def tau_ptr(self):
        return lib().b200qr_plan_tau_device(self._h)

    def apply_q(self, dA_ptr, ldm, dC_ptr, ldc, ncols, transpose, stream=0):
        """C <- Q^T C (transpose) or Q C with the reflectors of the matrix this plan just factored."""
        check(lib().b200qr_apply_q(C.c_void_p(dA_ptr), self.m, self.n, ldm, C.c_void_p(self.tau_ptr), C.c_void_p(dC_ptr),
                                   ldc, ncols, int(transpose), C.c_void_p(stream)), "b200qr_apply_q")

    def read_trace(self):
        import numpy as np
        n = self.tile_tasks
        out = np.zeros((n, 6), dtype=np.int32)
        check(lib().b200qr_plan_read_trace(self._h, out.ctypes.data_as(C.c_void_p), n), "b200qr_plan_read_trace")
        return out

    def close(self):
        if self._h:
            lib().b200qr_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TilePlan:
    """Tiled QR at tile size b in {32, 64, 128} (b200qr_tiles_*: the non-default mode of BASELINE.json's tile-size
    sweep; generic host-ordered kernels, own reflector layout = the reference's with 32 -> b)."""

    def __init__(self, m, n, b):
        self.m, self.n, self.b = m, n, b
        self._h = C.c_void_p()
        check(lib().b200qr_tiles_create(C.byref(self._h), m, n, b), "b200qr_tiles_create")

    def factor(self, dptr, ldm, stream=0):
        check(lib().b200qr_tiles_factor(self._h, C.c_void_p(dptr), ldm, C.c_void_p(stream)), "b200qr_tiles_factor")

    @property
    def tau_ptr(self):
        return lib().b200qr_tiles_tau(self._h)

    def apply_q(self, dA_ptr, ldm, dC_ptr, ldc, ncols, transpose, stream=0):
        check(lib().b200qr_tiles_apply_q(self._h, C.c_void_p(dA_ptr), ldm, C.c_void_p(dC_ptr), ldc, ncols, int(transpose),
                                         C.c_void_p(stream)), "b200qr_tiles_apply_q")

    def tau_reference_layout(self):
        """tau as the reference's tau matrix with 32 -> b (tile (i, k): rows 0..b-1 of the tile's first column)."""
        import numpy as np
        p, q, b = self.m // self.b, self.n // self.b, self.b
        K = min(p, q)
        tc = np.zeros((K, p, b), dtype=np.float32)
        check(lib().b200qr_copy(tc.ctypes.data_as(C.c_void_p), C.c_void_p(self.tau_ptr), tc.nbytes), "b200qr_copy")
        out = np.zeros((self.n, self.m), dtype=np.float32)
        for k in range(K):
            for i in range(k, p):
                out[k * b, i * b:(i + 1) * b] = tc[k, i]
        return out

    def close(self):
        if self._h:
            lib().b200qr_tiles_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AnyQR:
    """Arbitrary-size front-end (b200qr_any_*): m, n need not be multiples of 32."""

    def __init__(self, m, n, engine=0):
        self.m, self.n = m, n
        self._h = C.c_void_p()
        o = Opts()
        o.engine = int(engine)
        check(lib().b200qr_any_create(C.byref(self._h), m, n, C.byref(o)), "b200qr_any_create")

    def factor(self, dA_ptr, ldm, stream=0):
        check(lib().b200qr_any_factor(self._h, C.c_void_p(dA_ptr), ldm, C.c_void_p(stream)), "b200qr_any_factor")

    def get_r(self, dR_ptr, ldr, stream=0):
        check(lib().b200qr_any_get_r(self._h, C.c_void_p(dR_ptr), ldr, C.c_void_p(stream)), "b200qr_any_get_r")

    def apply_q(self, dC_ptr, ldc, ncols, transpose, stream=0):
        check(lib().b200qr_any_apply_q(self._h, C.c_void_p(dC_ptr), ldc, ncols, int(transpose), C.c_void_p(stream)),
              "b200qr_any_apply_q")

    def solve_ls(self, dB_ptr, ldb, nrhs, dX_ptr, ldx, stream=0):
        check(lib().b200qr_any_solve_ls(self._h, C.c_void_p(dB_ptr), ldb, nrhs, C.c_void_p(dX_ptr), ldx, C.c_void_p(stream)),
              "b200qr_any_solve_ls")

    def close(self):
        if self._h:
            lib().b200qr_any_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def factor_host(a, m, n, ldm=None, want_tau=False):
    """b200qr_factor_host on a NumPy (n, ldm) float32 array (column-major m x n), in place.
    Returns (kernel_ms, tau-or-None)."""
    import numpy as np
    ldm = ldm or m
    assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == (n, ldm)
    tau = np.zeros_like(a) if want_tau else None
    ms = C.c_float(0)
    check(lib().b200qr_factor_host(a.ctypes.data_as(C.c_void_p), m, n, ldm,
                                   tau.ctypes.data_as(C.c_void_p) if want_tau else None, C.byref(ms)),
          "b200qr_factor_host")
    return ms.value, tau


def sched_trace(M, N):
    """Serial trace of the host scheduler mirror with the reference's pick rule."""
    import numpy as np
    L = lib()
    s = L.b200qr_sched_init(M, N)
    t = Task()
    out = []
    while L.b200qr_sched_next(s, C.byref(t)) == TASK_AVAIL:
        out.append((t.type, t.l, t.m, t.k))
        L.b200qr_sched_done(s, C.byref(t))
    L.b200qr_sched_free(s)
    return np.array(out, dtype=np.int32).reshape(-1, 4)


def device_order(M, N):
    import numpy as np
    cap = sum((M - k) * (N - k) for k in range(min(M, N)))
    out = np.zeros((cap, 4), dtype=np.int32)
    n = lib().b200qr_sched_device_order(M, N, out.ctypes.data_as(C.c_void_p), cap)
    assert n == cap, (n, cap)
    return out


class DistRank:
    """One rank of the 1-D block-cyclic multi-GPU factorisation (b200qr_dist_*)."""

    def __init__(self, m, n, rank, world, share_device=0, grid_ctas=0, trace=0, engine=0):
        self.m, self.n, self.rank, self.world = m, n, rank, world
        self._h = C.c_void_p()
        o = Opts()
        o.grid_ctas, o.share_device, o.trace, o.engine = int(grid_ctas), int(share_device), int(trace), int(engine)
        check(lib().b200qr_dist_create(C.byref(self._h), m, n, rank, world, C.byref(o)), "b200qr_dist_create")
        self.block = lib().b200qr_dist_block(self._h)

    def export(self):
        buf = C.create_string_buffer(lib().b200qr_dist_handle_bytes())
        check(lib().b200qr_dist_export(self._h, buf), "b200qr_dist_export")
        return buf.raw

    def connect(self, blobs):
        """blobs: the exported handle bytes of all ranks, in rank order."""
        raw = b"".join(blobs)
        check(lib().b200qr_dist_connect(self._h, C.c_char_p(raw)), "b200qr_dist_connect")

    @property
    def matrix_ptr(self):
        return lib().b200qr_dist_matrix(self._h)

    @property
    def tau_ptr(self):
        return lib().b200qr_dist_tau_device(self._h)

    @property
    def grid(self):
        return lib().b200qr_dist_grid(self._h)

    @property
    def variant(self):
        return lib().b200qr_dist_variant(self._h)

    def owner(self, tile_col):
        return (tile_col // self.block) % self.world

    def owned_tile_cols(self):
        return [j for j in range(self.n // TILE) if self.owner(j) == self.rank]

    def upload(self, src_ptr):
        """Copy a full m x n float32 matrix (host or device pointer, ld = m) into the rank's matrix."""
        check(lib().b200qr_copy(C.c_void_p(self.matrix_ptr), C.c_void_p(src_ptr), self.m * self.n * 4), "b200qr_copy")

    def download(self, dst_ptr):
        check(lib().b200qr_copy(C.c_void_p(dst_ptr), C.c_void_p(self.matrix_ptr), self.m * self.n * 4), "b200qr_copy")

    def download_tau(self, dst_ptr):
        """Compact tau [(k*p + i)*32 + r], K*p*32 floats."""
        p, K = self.m // TILE, min(self.m, self.n) // TILE
        check(lib().b200qr_copy(C.c_void_p(dst_ptr), C.c_void_p(self.tau_ptr), K * p * TILE * 4), "b200qr_copy")

    def read_trace(self):
        import numpy as np
        n = lib().b200qr_dist_tile_tasks(self._h)
        out = np.zeros((n, 6), dtype=np.int32)
        check(lib().b200qr_dist_read_trace(self._h, out.ctypes.data_as(C.c_void_p), n), "b200qr_dist_read_trace")
        return out

    def reset(self, stream=0):
        check(lib().b200qr_dist_reset(self._h, C.c_void_p(stream)), "b200qr_dist_reset")

    def factor(self, stream=0):
        check(lib().b200qr_dist_factor(self._h, C.c_void_p(stream)), "b200qr_dist_factor")

    def reset_host(self):
        """Progress words = 'input not here yet' (synchronous); barrier, then factor_host."""
        check(lib().b200qr_dist_reset_host(self._h), "b200qr_dist_reset_host")

    def factor_host(self, host_ptr, ldh, packed=False):
        """Factor with HOST data (pinned), copies overlapped with the kernel; returns the kernel's milliseconds.
        packed: host_ptr holds only this rank's tile columns, consecutively; else the whole m x n matrix."""
        ms = C.c_float(0.0)
        check(lib().b200qr_dist_factor_host(self._h, C.c_void_p(host_ptr), C.c_longlong(ldh), int(bool(packed)), C.byref(ms)),
              "b200qr_dist_factor_host")
        return ms.value

    def close(self):
        if self._h:
            lib().b200qr_dist_destroy(self._h)
            self._h = C.c_void_p()


DEFAULT_TSQR_MODE = "stream"


class Tsqr:
    """Tall-skinny QR of an m x n matrix, n <= 256 (b200qr_tsqr_*).  Default: the streaming engine (strips of rows, the
    n x n triangle resident in shared memory); leaf_rows > 0 selects the leaf-batch tree (reference tile layout per leaf)."""

    def __init__(self, m, n, strips=0, leaf_rows=0, fanin=4, rstack=0, mode=None):
        """mode: "stream" (streaming engine), "tiles" (one warp per leaf on the batched tile kernel); None = the
        library default (QRB200_TSQR_MODE overrides)."""
        self.m, self.n = m, n
        self._h = C.c_void_p()
        mode = mode or os.environ.get("QRB200_TSQR_MODE") or DEFAULT_TSQR_MODE
        self.mode = "rstack" if rstack else ("leaf-batch tree" if leaf_rows else mode)
        if rstack:                  # m = rstack * n rows: a stack of `rstack` upper-triangular factors
            check(lib().b200qr_tsqr_create_rstack(C.byref(self._h), rstack, n), "b200qr_tsqr_create_rstack")
        elif not leaf_rows and mode == "tiles":
            check(lib().b200qr_tsqr_create_tiles(C.byref(self._h), m, n, strips), "b200qr_tsqr_create_tiles")
        elif leaf_rows:
            check(lib().b200qr_tsqr_create(C.byref(self._h), m, n, leaf_rows, fanin), "b200qr_tsqr_create")
        else:
            check(lib().b200qr_tsqr_create_streaming(C.byref(self._h), m, n, strips), "b200qr_tsqr_create_streaming")

    @property
    def levels(self):
        return lib().b200qr_tsqr_levels(self._h)

    @property
    def kernel_name(self):
        return "batched_warp_kernel" if self.mode == "tiles" else "tsqr_strip_kernel"

    def describe(self):
        if self.mode == "tiles":
            return "tile engine: %d leaves of %d rows, one warp each; %d pairwise levels" % (
                lib().b200qr_tsqr_leaves(self._h), self.m // max(1, lib().b200qr_tsqr_leaves(self._h)), self.levels - 1)
        return "%s engine, %d levels" % (self.mode, self.levels)

    def factor(self, dA_ptr, ldm, dR_ptr, stream=0):
        """dA (m x n, column-major, ldm) is overwritten with the leaf reflectors, dR (n x n, ld = n) receives R."""
        check(lib().b200qr_tsqr_factor(self._h, C.c_void_p(dA_ptr), ldm, C.c_void_p(dR_ptr), C.c_void_p(stream)),
              "b200qr_tsqr_factor")

    def close(self):
        if self._h:
            lib().b200qr_tsqr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tsqr_gather_stack(R_local, world, group=None):
    """The exchange step of the multi-GPU TSQR: all-gather the ranks' n x n R factors (column-major: a torch (n, n)
    tensor indexed [column, row]) and lay them out as the column-major (world*n) x n stack the top-level reduction reads,
    element (g*n + r, c) at [c, g*n + r].  NCCL over NVLink on GPUs; gloo in the CPU tests."""
    import torch
    import torch.distributed as dist
    n = R_local.shape[0]
    parts = [torch.empty_like(R_local) for _ in range(world)]
    dist.all_gather(parts, R_local.contiguous(), group=group)
    return torch.stack(parts, 0).permute(1, 0, 2).reshape(n, world * n).contiguous()


class TsqrDist:
    """BASELINE.json configs[3] on W GPUs: the rows are split into W contiguous blocks, rank r owns block r.
    Every rank reduces its block to an n x n R factor (streaming engine), the factors are all-gathered -- the path's one
    real exchange step, W x 256 KB at n = 256 -- and every rank reduces the stack of W factors redundantly, so all ranks
    end with the same R.  The implicit Q is the rank's leaf reflectors (in place in its block) plus the two trees.
    `local_reduce(A, R)` / `top_reduce(stack, R)` are the CUDA plans (there is no CPU implementation here); the gloo
    test of the exchange step passes its own stand-ins for them, since that test runs without a GPU."""

    def __init__(self, m_local, n, world, group=None, local_reduce=None, top_reduce=None):
        self.m_local, self.n, self.world, self.group = m_local, n, world, group
        self._local = self._top = None
        if local_reduce is None:
            self._local = Tsqr(m_local, n)
            self._top = Tsqr(world * n, n, rstack=world) if world > 1 else None
        self.local_reduce = local_reduce or self._cuda_local
        self.top_reduce = top_reduce or self._cuda_top

    def _cuda_local(self, A, R):
        import torch
        self._local.factor(A.data_ptr(), self.m_local, R.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def _cuda_top(self, stack, R):
        import torch
        self._top.factor(stack.data_ptr(), self.world * self.n, R.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def factor(self, A, R):
        """A: this rank's (n, m_local) tensor = column-major m_local x n block, overwritten with its reflectors;
        R: (n, n) tensor receiving the final factor (column-major, [column, row]).  Returns the local R factor."""
        import torch
        R_local = torch.empty_like(R)
        self.local_reduce(A, R_local)
        if self.world == 1:
            R.copy_(R_local)
            return R_local
        stack = tsqr_gather_stack(R_local, self.world, self.group)
        self.top_reduce(stack, R)
        self.stack = stack                      # the top tree's reflectors (kept: part of the implicit Q)
        return R_local

    def close(self):
        for t in (self._local, self._top):
            if t is not None:
                t.close()


class TsqrDistC:
    """One rank of the multi-GPU TSQR behind the C ABI (b200qr_tsqr_dist_*): the R factors are exchanged over peer
    memory from inside the stream, no collective library on the data path.  export() / connect(blobs) like DistRank."""

    def __init__(self, m_local, n, rank, world, mode=None):
        self.m_local, self.n, self.rank, self.world = m_local, n, rank, world
        mode = mode or os.environ.get("QRB200_TSQR_MODE") or DEFAULT_TSQR_MODE
        self.mode = mode
        self._h = C.c_void_p()
        check(lib().b200qr_tsqr_dist_create(C.byref(self._h), m_local, n, rank, world, 1 if mode == "tiles" else 0),
              "b200qr_tsqr_dist_create")

    def export(self):
        buf = C.create_string_buffer(lib().b200qr_tsqr_dist_handle_bytes())
        check(lib().b200qr_tsqr_dist_export(self._h, buf), "b200qr_tsqr_dist_export")
        return buf.raw

    def connect(self, blobs):
        check(lib().b200qr_tsqr_dist_connect(self._h, C.c_char_p(b"".join(blobs))), "b200qr_tsqr_dist_connect")

    def factor(self, dA_ptr, ldm, dR_ptr, stream=0):
        check(lib().b200qr_tsqr_dist_factor(self._h, C.c_void_p(dA_ptr), ldm, C.c_void_p(dR_ptr), C.c_void_p(stream)),
              "b200qr_tsqr_dist_factor")

    def factor_begin(self, dA_ptr, ldm, stream=0):
        check(lib().b200qr_tsqr_dist_factor_begin(self._h, C.c_void_p(dA_ptr), ldm, C.c_void_p(stream)), "b200qr_tsqr_dist_factor_begin")

    def factor_end(self, dR_ptr, stream=0):
        check(lib().b200qr_tsqr_dist_factor_end(self._h, C.c_void_p(dR_ptr), C.c_void_p(stream)), "b200qr_tsqr_dist_factor_end")

    def close(self):
        if self._h:
            lib().b200qr_tsqr_dist_destroy(self._h)
            self._h = C.c_void_p()


def tsqr_dist_connect_local(ranks):
    arr = (C.c_void_p * len(ranks))(*[r._h for r in ranks])
    check(lib().b200qr_tsqr_dist_connect_local(arr, len(ranks)), "b200qr_tsqr_dist_connect_local")


def dist_connect_local(ranks):
    arr = (C.c_void_p * len(ranks))(*[r._h for r in ranks])
    check(lib().b200qr_dist_connect_local(arr, len(ranks)), "b200qr_dist_connect_local")


def dist_factor_group(ranks, stream=0):
    arr = (C.c_void_p * len(ranks))(*[r._h for r in ranks])
    check(lib().b200qr_dist_factor_group(arr, len(ranks), C.c_void_p(stream)), "b200qr_dist_factor_group")


def rank_tickets(M, N, G, rank, world):
    """(sweeps [(k, g)...], panels [k...]) of one rank of the block-cyclic distribution."""
    import numpy as np
    L = lib()
    npan = C.c_long(0)
    ns = L.b200qr_sched_rank_tickets(M, N, G, rank, world, None, 0, None, 0, C.byref(npan))
    sw = np.zeros((max(ns, 1), 2), dtype=np.int32)
    pa = np.zeros(max(npan.value, 1), dtype=np.int32)
    L.b200qr_sched_rank_tickets(M, N, G, rank, world, sw.ctypes.data_as(C.c_void_p), ns,
                                pa.ctypes.data_as(C.c_void_p), npan.value, C.byref(npan))
    return sw[:ns], pa[:npan.value]
