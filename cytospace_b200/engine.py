"""Device-side driver of the assignment hot path: cost build -> LAP -> spot per cell.

PyTorch is used only for device memory (tensors as handles), streams and the
optional ``torch.distributed`` plumbing; every computation is a call through the
C ABI of ``libcytospace_b200.so`` (``include/cytospace_b200.h``).  There is no CPU
fallback: without a CUDA device / the built library every entry raises.

Reference path replaced (paths relative to the CytoSPACE repo):
``solve_linear_assignment_problem`` cytospace/cytospace.py:304-351 ->
``calculate_cost`` linear_assignment_solvers.py:42-69 ->
``matrix_correlation_pearson`` common/common.py:190-199 ->
``call_solver`` linear_assignment_solvers.py:34-40 (third-party ``lapjv``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _native

COST_SCALE = 10 ** 6           # integer scale of the correlation distance; precedent cytospace.py:337
PRECISIONS = {"f16": 0, "f16x3": 1}
#: ``--distance-metric`` values (argument_parser.py:72-74) -> CYB_METRIC_*
METRICS = {"Pearson_correlation": 0, "Spearman_correlation": 1, "Euclidean": 2}
#: kernels one C-ABI call launches (profiles/r02_ncu_launches_cfg2.csv lists them): colsum + colstat_finalize +
#: standardise_write per matrix and one GEMM for a Pearson cost build; one more column-sum pass per matrix with the
#: fused normalize_data; a rank kernel per matrix (and the statistics passes of the ranks) for Spearman
KERNELS_PER_CALL = {"cost_pearson": 7, "cost_log_tpm_extra": 2, "cost_spearman_extra": 2, "lap_solve": 1,
                    "lap_check_whole": 1, "lap_check_tiled": 3, "rank_columns": 1, "expand": 1, "quantise": 1}
STAT_NAMES = ("status", "phases", "rounds", "bids", "passes", "cost_min", "cost_max", "scale",
              "grid", "smem_prices", "tail_mode", "max_bidders", "phase_scans", "tail_bids", "tails", "list_hits",
              "small_rounds", "ns_bid", "ns_barrier", "ns_resolve", "ns_tail", "paths", "ns_select", "ns_relax",
              "ns_augment", "ns_sel_pass", "ns_sel_scan", "warm",
              "ns_phase_start", "ns_auction", "ns_init", "ns_total")
#: with the default solver (auction rounds + shortest-augmenting-path finish, csrc/lap_sap.cu) the tail
#: columns read: tail_bids = rows relaxed in searches, tails = searches, list_hits = search rounds


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def _on_engine_device(fn):
    """Run a method with the engine's device current: the native library queries and launches on the
    current CUDA device."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapper


@dataclass
class LapResult:
    """Device-resident LAP output.  With unit capacities ``person_obj`` / ``slot_owner`` are
    lapjv's ``row_ind`` / ``col_ind``; CytoSPACE consumes ``col_ind``
    (linear_assignment_solvers.py:38: the row assigned to each column)."""
    person_obj: torch.Tensor  # int32[P]  object (spot) of person (cell) i
    slot_owner: torch.Tensor  # int32[P]  person in slot t (slots ordered by object)
    price: torch.Tensor       # int64[O]  object prices, units of 1/(P+1) cost
    total: int                # sum_i cost[i, person_obj[i]]
    stats: dict
    slot_offsets: torch.Tensor | None = None   # int32[O+1] device (None: unit capacities)

    @property
    def rowsol(self):
        return self.person_obj

    @property
    def colsol(self):
        return self.slot_owner

    @property
    def row_scans(self) -> int:
        """Row scans the solve performed (round bids + tail bids + phase-start re-checks)."""
        return int(self.stats["bids"]) + int(self.stats["tail_bids"]) + int(self.stats["phase_scans"])


class AssignmentEngine:
    """One engine per (process, device).  Buffers are cached and re-used between calls."""

    def __init__(self, device=None, precision: str = "f16x3", cost_scale: float = COST_SCALE):
        if not torch.cuda.is_available():
            raise RuntimeError("cytospace_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _native.load()
        self.ffi = _native.ffi()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        self.cost_scale = float(cost_scale)
        self._ws = {}
        #: kernels of this library launched so far (every C-ABI call launches a fixed number: see KERNELS_PER_CALL)
        self.launches = 0
        self.h2d_bytes = 0                 # bytes to_device() has sent over PCIe (the device tensors' sizes)
        self.profile = False          # True: bracket the cost-build and LAP launches with CUDA events
        self._events = {}
        sm, maj, mnr, mem = (self.ffi.new("int *"), self.ffi.new("int *"), self.ffi.new("int *"),
                             self.ffi.new("size_t *"))
        _native.check(self.lib.cyb_device_info(self.device.index, sm, maj, mnr, mem))
        self.sm_count, self.cc, self.total_mem = sm[0], (maj[0], mnr[0]), mem[0]
        if self.cc[0] != 10:
            raise RuntimeError(f"cytospace_b200 is built for sm_100a; device is cc {self.cc[0]}.{self.cc[1]}")

    # ------------------------------------------------------------------ helpers
    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            self._ws[key] = None
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    def _stream(self):
        return _native.stream_ptr(torch.cuda.current_stream(self.device))

    def _mark(self, name: str, which: int):
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self._events.setdefault(name, [None, None])[which] = ev

    def last_ms(self, name: str) -> float:
        """Device time of the last ``cost`` / ``lap`` launch group (needs ``profile = True``)."""
        a, b = self._events[name]
        b.synchronize()
        return a.elapsed_time(b)

    #: host arrays at least this large go through the pinned staging ring
    STAGE_MIN_BYTES = 32 << 20
    #: large float64 host arrays cross PCIe as float32 (narrowed by the staging threads; half the bytes).  Exact for count
    #: matrices; on normalised data the correlation moves by < 1e-8 (include/cytospace_b200.h).  ``CYB_STAGE_F32=0`` or
    #: ``engine.stage_float32 = False`` keeps float64 on the wire.
    stage_float32 = os.environ.get("CYB_STAGE_F32", "1") != "0"
    last_stage_exact = True

    @_on_engine_device
    def to_device(self, x, dtype=None) -> torch.Tensor:
        """Host array (numpy / CPU tensor) -> device tensor.

        Expression matrices reach ``solve_linear_assignment_problem`` as pageable numpy arrays
        (cytospace.py:398-409).  Anything that is not float32 / float64 (``read_csv`` count matrices are
        int64) is cast to float64 on the host as the reference's ``normalize_data`` does
        (common.py:143: ``np.nan_to_num(data).astype(float)``).  Large float64 arrays land as float32 unless
        ``stage_float32`` is off or ``dtype`` is given.  Large arrays are uploaded through the native library's ring
        of pinned pieces (``cyb_stage_upload``; a plain ``cudaMemcpy`` from pageable memory runs at a fraction of the
        PCIe rate)."""
        if torch.is_tensor(x):
            if x.is_cuda:
                return x if dtype is None or x.dtype == dtype else x.to(dtype)
            x = x.numpy()
        x = np.asarray(x)
        if dtype is not None:
            want = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32,
                    torch.int64: np.int64}[dtype]
            if x.dtype != want:
                x = x.astype(want)
        elif x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        if not x.flags.c_contiguous:
            x = np.ascontiguousarray(x)
        if x.nbytes < self.STAGE_MIN_BYTES:
            self.h2d_bytes += x.nbytes
            return torch.from_numpy(x).to(self.device)
        out = self._staged_upload(x, narrow=(dtype is None and x.dtype == np.float64 and self.stage_float32))
        self.h2d_bytes += out.numel() * out.element_size()
        return out

    def _staged_upload(self, x: np.ndarray, narrow: bool = False) -> torch.Tensor:
        """``cyb_stage_upload``: worker threads of the native library copy pageable -> pinned pieces with
        non-temporal stores and enqueue one DMA per piece; returns when the last piece is enqueued.
        ``narrow``: a float64 array arrives as float32 (``cyb_stage_upload_f64_as_f32``; ``last_stage_exact`` tells
        whether every value survived exactly)."""
        if narrow:
            out = torch.empty(x.shape, dtype=torch.float32, device=self.device)
            flag = _native.ffi().new("int32_t *")
            _native.check(self.lib.cyb_stage_upload_f64_as_f32(_native.ffi().cast("const double *", x.ctypes.data),
                                                               _native.ptr("float *", out), x.size, flag, self._stream()))
            self.last_stage_exact = not bool(flag[0])
            return out
        out = torch.empty(x.shape, dtype=torch.from_numpy(x[:0]).dtype, device=self.device)
        _native.check(self.lib.cyb_stage_upload(_native.ffi().cast("const void *", x.ctypes.data),
                                                _native.ptr("void *", out), x.nbytes, self._stream()))
        return out

    @_on_engine_device
    def narrow_exact(self, x: torch.Tensor):
        """float32 copy of a contiguous float64 device matrix if EVERY value is exactly representable (raw counts are),
        else None -- ``cyb_narrow_f64_to_f32``, one pass.  Used for the wire format of chunk blocks."""
        if x.dtype != torch.float64 or not x.is_contiguous():
            return None
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
        _native.check(self.lib.cyb_narrow_f64_to_f32(_native.ptr("double *", x), x.numel(), _native.ptr("float *", out),
                                                     _native.ptr("int32_t *", flag), self._stream()))
        self.launches += 1
        return None if int(flag.item()) else out

    @_on_engine_device
    def cost_build(self, sc: torch.Tensor, st: torch.Tensor, log_tpm: bool = False, out: torch.Tensor | None = None,
                   return_colstats: bool = False, check_variance: bool = True, layout: str = "cells_x_spots",
                   metric: str = "Pearson_correlation"):
        """Integer cost as int32, row-major with ld = columns rounded up to 32:
        ``rint(-cost_scale * r)`` for the Pearson / Spearman correlation r, ``rint(cost_scale * d)`` for
        the Euclidean distance d (linear_assignment_solvers.py:46-59).
        ``layout="cells_x_spots"`` ([N, ld]: the TRANSPOSE of the reference's ``cost``, a
        cell's row contiguous) or ``"spots_x_cells"`` ([S, ld]: the reference's own orientation).

        sc [G x N], st [G x S]: float64 / float32 device tensors, genes x cells row-major -- the
        arrays ``calculate_cost`` receives (linear_assignment_solvers.py:42).  Raises ``ValueError``
        when the gene counts differ (common.py:191-192) or (documented deviation; the reference
        yields NaN, common.py:196-197) when a column has zero variance."""
        if sc.dim() != 2 or st.dim() != 2:
            raise ValueError("expression matrices must be 2-D (genes x cells)")
        if sc.shape[0] != st.shape[0]:
            raise ValueError("The two matrices v1 and v2 must have equal dimensions; "
                             "ST and scRNA data must have the same genes")
        if sc.dtype != st.dtype or sc.dtype not in (torch.float64, torch.float32):
            raise ValueError("expression matrices must both be float64 or both float32")
        if not (sc.is_cuda and st.is_cuda):
            raise ValueError("cost_build takes device tensors (use to_device())")
        if sc.stride(1) != 1:
            sc = sc.contiguous()
        if st.stride(1) != 1:
            st = st.contiguous()
        G, N = sc.shape
        S = st.shape[1]
        if G == 0 or N == 0 or S == 0:
            raise ValueError("empty expression matrix")
        prec = PRECISIONS[self.precision]
        if metric not in METRICS:
            raise ValueError(f"distance_metric must be one of {sorted(METRICS)}")
        if return_colstats and metric != "Pearson_correlation":
            raise ValueError("column statistics are only returned for Pearson_correlation")
        if layout not in ("cells_x_spots", "spots_x_cells"):
            raise ValueError("layout must be 'cells_x_spots' or 'spots_x_cells'")
        rows, cols = (N, S) if layout == "cells_x_spots" else (S, N)
        ld = _round_up(cols, 32)
        if out is None:
            out = torch.empty((rows, ld), dtype=torch.int32, device=self.device)
        elif out.shape[0] < rows or out.stride(0) < cols or out.dtype != torch.int32:
            raise ValueError("bad `out` buffer")
        ws_bytes = self.lib.cyb_cost_build_metric_workspace_bytes(METRICS[metric], G, N, S, prec)
        ws = self._workspace("cost", ws_bytes + 1024)
        off = (-ws.data_ptr()) % 1024
        colstat_sc = torch.empty((2, N), dtype=torch.float64, device=self.device) if return_colstats else None
        colstat_st = torch.empty((2, S), dtype=torch.float64, device=self.device) if return_colstats else None
        zero_var = torch.zeros(1, dtype=torch.int32, device=self.device)
        dt = self.lib.CYB_F64 if sc.dtype == torch.float64 else self.lib.CYB_F32
        self._mark("cost", 0)
        # the library builds cost[first, second]; the first matrix supplies the rows
        a, b, na, nb, csa, csb = (sc, st, N, S, colstat_sc, colstat_st) if layout == "cells_x_spots" \
            else (st, sc, S, N, colstat_st, colstat_sc)
        if metric == "Pearson_correlation":
            _native.check(self.lib.cyb_cost_build_pearson(
                _native.ptr("void *", a), _native.ptr("void *", b), dt, G, na, nb, a.stride(0), b.stride(0),
                int(bool(log_tpm)), prec, self.cost_scale, _native.ptr("int32_t *", out), out.stride(0),
                _native.ptr("double *", csa), _native.ptr("double *", csb),
                _native.ptr("int32_t *", zero_var), self.ffi.cast("void *", ws.data_ptr() + off), ws_bytes,
                self._stream()))
        else:
            _native.check(self.lib.cyb_cost_build(
                METRICS[metric], _native.ptr("void *", a), _native.ptr("void *", b), dt, G, na, nb, a.stride(0),
                b.stride(0), int(bool(log_tpm)), prec, self.cost_scale, _native.ptr("int32_t *", out), out.stride(0),
                _native.ptr("int32_t *", zero_var), self.ffi.cast("void *", ws.data_ptr() + off), ws_bytes,
                self._stream()))
        self._mark("cost", 1)
        self.launches += (KERNELS_PER_CALL["cost_pearson"] + (KERNELS_PER_CALL["cost_log_tpm_extra"] if log_tpm else 0) +
                          (KERNELS_PER_CALL["cost_spearman_extra"] if metric == "Spearman_correlation" else 0))
        self._zero_var = zero_var
        self._zero_var_metric = metric
        if check_variance:
            self.check_zero_variance()
        if return_colstats:
            return out, colstat_sc, colstat_st
        return out

    def check_zero_variance(self):
        """Raises if the last cost build met a zero-variance column (one small D2H read)."""
        nz = int(self._zero_var.item())
        if nz and getattr(self, "_zero_var_metric", "") == "Euclidean":
            raise ValueError(f"{nz} distance(s) cannot be represented (cost_scale * distance >= 2^30)")
        if nz:
            raise ValueError(f"{nz} cell/spot column(s) have zero variance: correlation undefined "
                             "(the reference would hand NaN costs to the solver)")

    @_on_engine_device
    def rank_columns(self, x: torch.Tensor, log_tpm: bool = False) -> torch.Tensor:
        """``pd.DataFrame(x).rank().values`` (common.py:207-208) for a [G x n] float64 / float32 device
        matrix: float32 [G x n] average ranks."""
        if x.dim() != 2 or not x.is_cuda or x.dtype not in (torch.float64, torch.float32):
            raise ValueError("rank_columns takes a 2-D float64/float32 device matrix")
        if x.stride(1) != 1:
            x = x.contiguous()
        G, n = x.shape
        out = torch.empty((G, n), dtype=torch.float32, device=self.device)
        ws_bytes = self.lib.cyb_rank_workspace_bytes(G, n)
        ws = self._workspace("rank", ws_bytes + 256)
        off = (-ws.data_ptr()) % 256
        dt = self.lib.CYB_F64 if x.dtype == torch.float64 else self.lib.CYB_F32
        self._mark("rank", 0)
        _native.check(self.lib.cyb_rank_columns(_native.ptr("void *", x), dt, G, n, x.stride(0), int(bool(log_tpm)),
                                                _native.ptr("float *", out), out.stride(0),
                                                self.ffi.cast("void *", ws.data_ptr() + off), ws_bytes, self._stream()))
        self._mark("rank", 1)
        self.launches += KERNELS_PER_CALL["rank_columns"]
        return out

    @_on_engine_device
    def expand_with_noise(self, cost: torch.Tensor, n_cols: int, row_map, seed: int, noise_lo: int = 1,
                          noise_span: int = 10) -> torch.Tensor:
        """``cost[location_repeat, :]`` (linear_assignment_solvers.py:63-66) plus the integer tie noise
        of the lap_CSPR path (cytospace.py:337-340): int32 [n_slots x ld]."""
        rm = torch.as_tensor(np.asarray(row_map), dtype=torch.int32).to(self.device)
        n_rows = int(rm.numel())
        ld = _round_up(n_cols, 32)
        out = torch.empty((n_rows, ld), dtype=torch.int32, device=self.device)
        _native.check(self.lib.cyb_expand_rows_noise_i32(
            _native.ptr("int32_t *", cost), cost.stride(0), n_rows, n_cols, _native.ptr("int32_t *", rm),
            int(seed) & ((1 << 64) - 1), int(noise_lo), int(noise_span), _native.ptr("int32_t *", out), ld,
            self._stream()))
        return out

    @_on_engine_device
    def quantise(self, cost_f64: torch.Tensor, scale: float) -> torch.Tensor:
        """int32 ``rint(scale * cost)`` of a float64 device matrix (entry P2)."""
        n_rows, n_cols = cost_f64.shape
        ld = _round_up(n_cols, 32)
        out = torch.empty((n_rows, ld), dtype=torch.int32, device=self.device)
        bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        _native.check(self.lib.cyb_quantise_f64(_native.ptr("double *", cost_f64), n_rows, n_cols, cost_f64.stride(0),
                                                float(scale), _native.ptr("int32_t *", out), ld,
                                                _native.ptr("int32_t *", bad), self._stream()))
        if int(bad.item()):
            raise ValueError("cost matrix contains NaN/inf or values too large to integerise")
        return out

    # ---------------------------------------------------------------------- LAP
    def _slot_offsets(self, capacities, n_objects):
        if capacities is None:
            return None, n_objects
        cap = np.asarray(capacities).astype(np.int64).ravel()
        if cap.shape[0] != n_objects:
            raise ValueError(f"{cap.shape[0]} capacities for {n_objects} objects")
        if (cap < 0).any():
            raise ValueError("negative capacity")
        if (cap == 1).all():
            return None, n_objects
        soff = np.zeros(n_objects + 1, dtype=np.int32)
        np.cumsum(cap, out=soff[1:])
        return torch.from_numpy(soff).to(self.device), int(soff[-1])

    @_on_engine_device
    def lap_solve(self, cost: torch.Tensor, capacities=None, n_persons: int | None = None,
                  n_objects: int | None = None, grid: int = 0) -> LapResult:
        """Exact assignment on the int32 device matrix ``cost[person, object]``: every person gets one
        object, object ``o`` exactly ``capacities[o]`` persons (None: 1 each, the square LAP)."""
        if cost.dtype != torch.int32 or not cost.is_cuda or cost.dim() != 2 or cost.stride(1) != 1:
            raise ValueError("cost must be a row-major int32 device matrix")
        if n_persons is None:
            n_persons = int(cost.shape[0])
        if n_objects is None:
            n_objects = n_persons if capacities is None else int(np.asarray(capacities).size)
        if n_persons <= 0 or n_objects <= 0:
            raise ValueError("empty LAP")
        if cost.shape[0] < n_persons or cost.shape[1] < n_objects:
            raise ValueError("LAP must be square: cost is smaller than persons x objects")
        soff, n_slots = self._slot_offsets(capacities, n_objects)
        if n_slots != n_persons:
            raise ValueError(f"the assignment must be square: {n_slots} slots for {n_persons} persons")
        dev = self.device
        person_obj = torch.empty(n_persons, dtype=torch.int32, device=dev)
        slot_owner = torch.empty(n_persons, dtype=torch.int32, device=dev)
        price = torch.empty(n_objects, dtype=torch.int64, device=dev)
        small = torch.zeros(1 + self.lib.CYB_LAP_NSTATS, dtype=torch.int64, device=dev)
        ws_bytes = self.lib.cyb_lap_workspace_bytes(n_persons, n_objects)
        ws = self._workspace("lap", ws_bytes + 256)
        off = (-ws.data_ptr()) % 256
        self._mark("lap", 0)
        _native.check(self.lib.cyb_lap_solve_i32(
            _native.ptr("int32_t *", cost), cost.stride(0), n_persons, n_objects, _native.ptr("int32_t *", soff),
            _native.ptr("int32_t *", person_obj), _native.ptr("int32_t *", slot_owner), _native.ptr("int64_t *", price),
            self.ffi.cast("int64_t *", small.data_ptr()), self.ffi.cast("int64_t *", small.data_ptr() + 8),
            self.ffi.cast("void *", ws.data_ptr() + off), ws_bytes, int(grid), self._stream()))
        self._mark("lap", 1)
        self.launches += KERNELS_PER_CALL["lap_solve"]
        host = small.cpu().tolist()            # one D2H: total + stats (synchronises the stream)
        stats = dict(zip(STAT_NAMES, host[1:1 + len(STAT_NAMES)]))
        if stats["status"] != 0:
            names = {self.lib.CYB_ERR_OVERFLOW: "price overflow", self.lib.CYB_ERR_NOT_CONVERGED: "round cap hit"}
            raise RuntimeError(f"LAP solve failed on device: {names.get(stats['status'], stats['status'])}")
        return LapResult(person_obj, slot_owner, price, int(host[0]), stats, soff)

    @_on_engine_device
    def lap_check(self, cost: torch.Tensor, res: LapResult) -> dict:
        """On-device optimality certificate; ``max_violation <= 1`` (scaled units) with no invalid
        person / capacity mismatch proves the assignment optimal for the integer matrix.  One
        coalesced pass over the matrix."""
        n_persons, n_objects = int(res.person_obj.numel()), int(res.price.numel())
        out = torch.zeros(4, dtype=torch.int64, device=self.device)
        ws_bytes = self.lib.cyb_lap_workspace_bytes(n_persons, n_objects)
        ws = self._workspace("lap", ws_bytes + 256)
        off = (-ws.data_ptr()) % 256
        self._mark("check", 0)
        _native.check(self.lib.cyb_lap_check_i32(
            _native.ptr("int32_t *", cost), cost.stride(0), n_persons, n_objects,
            _native.ptr("int32_t *", res.slot_offsets), _native.ptr("int32_t *", res.person_obj),
            _native.ptr("int64_t *", res.price), _native.ptr("int64_t *", out),
            self.ffi.cast("void *", ws.data_ptr() + off), ws_bytes, self._stream()))
        self._mark("check", 1)
        self.launches += KERNELS_PER_CALL["lap_check_whole" if n_objects <= 12288 else "lap_check_tiled"]
        v, t, bad, badcap = out.cpu().tolist()
        return {"max_violation": v, "total": t, "invalid_rows": bad, "capacity_mismatch": badcap}

    # --------------------------------------------------------------- whole path
    @_on_engine_device
    def assign(self, sc, st, cell_number_to_node_assignment, log_tpm: bool = False,
               metric: str = "Pearson_correlation", cspr_seed: int | None = None, progress=None):
        """cost build + LAP + ``location_repeat[assignment]`` (cytospace.py:319-331).

        ``cspr_seed`` not None selects the integerised lap_CSPR formulation (cytospace.py:334-347):
        the slot expansion is materialised with per-(slot, cell) integer noise in [1, 10] and solved
        as a square LAP (slots bid for cells).

        ``progress`` (a callable taking one string) receives the reference's progress lines
        (cytospace.py:321-322) as the corresponding stage is reached.

        Returns ``(spot_of_cell int64 device tensor [N], LapResult, cost int32 device matrix)``; the
        matrix is persons x objects of the solve: spots x cells when every spot takes one cell, cells x
        spots otherwise."""
        say = progress or (lambda _msg: None)
        say("Building cost matrix ...")
        cn = np.asarray(cell_number_to_node_assignment).astype(np.int64).ravel()
        sc = self.to_device(sc) if not (torch.is_tensor(sc) and sc.is_cuda) else sc
        st = self.to_device(st) if not (torch.is_tensor(st) and st.is_cuda) else st
        if sc.dtype != st.dtype or sc.dtype not in (torch.float64, torch.float32):
            # e.g. int64 counts on one side, float32 on the other: compute in float64 like the reference
            sc, st = sc.to(torch.float64), st.to(torch.float64)
        N, S = int(sc.shape[1]), int(st.shape[1])
        if cn.shape[0] != S:
            raise ValueError(f"cell_number_to_node_assignment has {cn.shape[0]} entries for {S} spots")
        if (cn < 0).any():
            raise ValueError("negative cell count")
        n = int(cn.sum())
        if n != N:
            raise ValueError(f"the assignment must be square: sum(cell_number_to_node_assignment)={n} "
                             f"but {N} cells were given")
        if cspr_seed is not None:
            compact = self.cost_build(sc, st, log_tpm=log_tpm, check_variance=False, layout="spots_x_cells",
                                      metric=metric)
            location_repeat = np.repeat(np.arange(S), cn)
            cost = self.expand_with_noise(compact, N, location_repeat, cspr_seed)
            say("Solving linear assignment problem ...")
            res = self.lap_solve(cost, None, n_persons=N, n_objects=N)
            lr = torch.from_numpy(location_repeat).to(self.device)
            spot_of_cell = lr[res.slot_owner.long()]
        elif (cn == 1).all():
            # square LAP: the spots bid for the cells (the reference's own orientation).  Either side
            # may bid; measured on B200 the noisier side (the cells) makes the better OBJECTS -- larger
            # gaps between a bidder's best and second-best object, shorter price wars.
            cost = self.cost_build(sc, st, log_tpm=log_tpm, check_variance=False, layout="spots_x_cells",
                                   metric=metric)
            say("Solving linear assignment problem ...")
            res = self.lap_solve(cost, None, n_persons=S, n_objects=N)
            spot_of_cell = res.slot_owner.long()
        else:
            # location_repeat (linear_assignment_solvers.py:63-65) becomes the spots' capacities:
            # the cells bid for spots that hold cn[s] cells each.  Spots that take no cell
            # (--sampling-sub-spots hands bincount(minlength=n_spots) vectors, cytospace.py:650-660) never
            # enter location_repeat in the reference; here they are dropped before the cost build, which also
            # keeps objects <= persons when there are more spots than cells in a chunk.
            keep = np.flatnonzero(cn > 0)
            if keep.size < S:
                st = st.index_select(1, torch.from_numpy(keep).to(self.device))
            cost = self.cost_build(sc, st, log_tpm=log_tpm, check_variance=False, layout="cells_x_spots",
                                   metric=metric)
            say("Solving linear assignment problem ...")
            res = self.lap_solve(cost, cn[keep], n_persons=N, n_objects=int(keep.size))
            spot_of_cell = res.person_obj.long()
            if keep.size < S:
                spot_of_cell = torch.from_numpy(keep).to(self.device)[spot_of_cell]
        self.check_zero_variance()                   # one sync, after the solve
        return spot_of_cell, res, cost
