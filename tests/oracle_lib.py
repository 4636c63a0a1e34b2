"""ctypes bindings for the parity oracle (TEST INFRASTRUCTURE).

* ``oracle/liboracle.so``            -- our plain-C restatement (oracle/emat_oracle.c)
* ``oracle/_ref/libdelphy_ref.so``   -- the reference's own sources compiled in place (oracle/ref_capi.cpp)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)


class OrcEmat(C.Structure):
    _fields_ = [
        ("num_nodes", C.c_int32), ("root", C.c_int32), ("includes_run_root", C.c_int32), ("reserved", C.c_int32),
        ("parent", i32p), ("child0", i32p), ("child1", i32p), ("t", f64p),
        ("mut_off", i32p), ("mut_site", i32p), ("mut_from", u8p), ("mut_to", u8p), ("mut_t", f64p),
        ("miss_off", i32p), ("miss_start", i32p), ("miss_end", i32p),
        ("fs_off", i32p), ("fs_site", i32p), ("fs_from", u8p),
    ]


class OrcSites(C.Structure):
    _fields_ = [
        ("num_sites", C.c_int32), ("num_partitions", C.c_int32),
        ("ref", u8p), ("partition_for_site", i32p), ("nu_l", f64p),
        ("mu", f64p), ("pi_a", f64p), ("q_ab", f64p),
    ]


class OrcRegion(C.Structure):
    _fields_ = [
        ("branch", C.c_int32), ("mut_idx", C.c_int32), ("t_min", C.c_double), ("t_max", C.c_double),
        ("min_muts", C.c_int32), ("pad_", C.c_int32), ("log_W_over_Wmax", C.c_double), ("W_over_Wmax", C.c_double),
    ]


class OrcStudySummary(C.Structure):
    _fields_ = [
        ("mu", C.c_double), ("log_Wmax", C.c_double), ("sum_W_over_Wmax", C.c_double),
        ("num_regions", C.c_int32), ("num_missing_at_X", C.c_int32),
    ]


REGION_DTYPE = np.dtype([
    ("branch", "<i4"), ("mut_idx", "<i4"), ("t_min", "<f8"), ("t_max", "<f8"),
    ("min_muts", "<i4"), ("pad_", "<i4"), ("log_W_over_Wmax", "<f8"), ("W_over_Wmax", "<f8"),
])
assert REGION_DTYPE.itemsize == 48 == C.sizeof(OrcRegion)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


@dataclass
class Emat:
    """One EMAT flattened to SoA+CSR in host node order (same layout as include/delphy_b200.h)."""
    root: int
    parent: np.ndarray
    child0: np.ndarray
    child1: np.ndarray
    t: np.ndarray
    mut_off: np.ndarray
    mut_site: np.ndarray
    mut_from: np.ndarray
    mut_to: np.ndarray
    mut_t: np.ndarray
    miss_off: np.ndarray
    miss_start: np.ndarray
    miss_end: np.ndarray
    fs_off: np.ndarray
    fs_site: np.ndarray
    fs_from: np.ndarray
    includes_run_root: int = 1
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        def a(x, dt):
            return np.ascontiguousarray(np.asarray(x, dtype=dt))
        self.parent = a(self.parent, np.int32); self.child0 = a(self.child0, np.int32); self.child1 = a(self.child1, np.int32)
        self.t = a(self.t, np.float64)
        self.mut_off = a(self.mut_off, np.int32); self.mut_site = a(self.mut_site, np.int32)
        self.mut_from = a(self.mut_from, np.uint8); self.mut_to = a(self.mut_to, np.uint8); self.mut_t = a(self.mut_t, np.float64)
        self.miss_off = a(self.miss_off, np.int32); self.miss_start = a(self.miss_start, np.int32); self.miss_end = a(self.miss_end, np.int32)
        self.fs_off = a(self.fs_off, np.int32); self.fs_site = a(self.fs_site, np.int32); self.fs_from = a(self.fs_from, np.uint8)

    @property
    def num_nodes(self) -> int:
        return int(self.parent.shape[0])

    def num_muts_of(self, v: int) -> int:
        return int(self.mut_off[v + 1] - self.mut_off[v])

    def as_struct(self) -> OrcEmat:
        return OrcEmat(
            self.num_nodes, int(self.root), int(self.includes_run_root), 0,
            _p(self.parent, i32p), _p(self.child0, i32p), _p(self.child1, i32p), _p(self.t, f64p),
            _p(self.mut_off, i32p), _p(self.mut_site, i32p), _p(self.mut_from, u8p), _p(self.mut_to, u8p), _p(self.mut_t, f64p),
            _p(self.miss_off, i32p), _p(self.miss_start, i32p), _p(self.miss_end, i32p),
            _p(self.fs_off, i32p), _p(self.fs_site, i32p), _p(self.fs_from, u8p))


@dataclass
class Sites:
    ref: np.ndarray
    partition_for_site: np.ndarray
    nu_l: np.ndarray
    mu: np.ndarray
    pi_a: np.ndarray
    q_ab: np.ndarray

    def __post_init__(self):
        def a(x, dt):
            return np.ascontiguousarray(np.asarray(x, dtype=dt))
        self.ref = a(self.ref, np.uint8); self.partition_for_site = a(self.partition_for_site, np.int32)
        self.nu_l = a(self.nu_l, np.float64); self.mu = a(self.mu, np.float64).reshape(-1)
        self.pi_a = a(self.pi_a, np.float64).reshape(-1, 4); self.q_ab = a(self.q_ab, np.float64).reshape(-1, 4, 4)

    @property
    def num_sites(self) -> int:
        return int(self.ref.shape[0])

    @property
    def num_partitions(self) -> int:
        return int(self.mu.shape[0])

    def as_struct(self) -> OrcSites:
        return OrcSites(self.num_sites, self.num_partitions, _p(self.ref, u8p), _p(self.partition_for_site, i32p),
                        _p(self.nu_l, f64p), _p(self.mu, f64p), _p(self.pi_a, f64p), _p(self.q_ab, f64p))


def emat_from_lists(root, parent, children, t, mutations, miss_intervals, from_states, includes_run_root=1) -> Emat:
    """mutations[v] = [(from, site, to, t), ...]; miss_intervals[v] = [(start, end), ...]; from_states[v] = [(site, from), ...]"""
    n = len(parent)
    mo, io, fo = [0], [0], [0]
    ms, mf, mt, mtt, ist, ien, fs, ff = [], [], [], [], [], [], [], []
    for v in range(n):
        for (fr, site, to, tt) in mutations[v]:
            ms.append(site); mf.append(fr); mt.append(to); mtt.append(tt)
        mo.append(len(ms))
        for (s, e) in sorted(miss_intervals[v]):
            ist.append(s); ien.append(e)
        io.append(len(ist))
        for (site, fr) in sorted(from_states[v]):
            fs.append(site); ff.append(fr)
        fo.append(len(fs))
    c0 = [c[0] if len(c) == 2 else -1 for c in children]
    c1 = [c[1] if len(c) == 2 else -1 for c in children]
    return Emat(root, parent, c0, c1, t, mo, ms, mf, mt, mtt, io, ist, ien, fo, fs, ff, includes_run_root)


class OrcApiTreeView(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("root", C.c_int32), ("num_sites", C.c_int32), ("pad_", C.c_int32),
                ("num_muts", C.c_int64), ("num_ivls", C.c_int64),
                ("nodes", C.c_void_p), ("muts", C.c_void_p), ("ivls", C.c_void_p), ("ref_seq", C.c_void_p)]


# parent, child0, child1, t, mut_off, mut_site, mut_from, mut_to, mut_t, miss_off, miss_start, miss_end, fs_off, fs_site, fs_from
_EMAT_OUT_ARGS = [i32p, i32p, i32p, f64p, i32p, i32p, u8p, u8p, f64p, i32p, i32p, i32p, i32p, i32p, u8p]


def _alloc_emat_arrays(counts):
    n, _, M, I, F = (int(c) for c in counts)
    return [np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64),
            np.zeros(n + 1, np.int32), np.zeros(max(M, 1), np.int32), np.zeros(max(M, 1), np.uint8), np.zeros(max(M, 1), np.uint8),
            np.zeros(max(M, 1), np.float64),
            np.zeros(n + 1, np.int32), np.zeros(max(I, 1), np.int32), np.zeros(max(I, 1), np.int32),
            np.zeros(n + 1, np.int32), np.zeros(max(F, 1), np.int32), np.zeros(max(F, 1), np.uint8)]


def _emat_from_arrays(counts, arrs) -> "Emat":
    n, root, M, I, F = (int(c) for c in counts)
    a = list(arrs)
    for i in (5, 6, 7, 8): a[i] = a[i][:M]
    for i in (10, 11): a[i] = a[i][:I]
    for i in (13, 14): a[i] = a[i][:F]
    return Emat(root, *a)


def _ptrs(arrs):
    tys = _EMAT_OUT_ARGS
    return [_p(a, ty) for a, ty in zip(arrs, tys)]


def api_tree_write(e: "Emat", ref_seq: np.ndarray, impl: str = "oracle", sites: "Sites | None" = None) -> bytes:
    """phylo_tree_to_api_tree (core/api.cpp:34-98): the size-prefixed FlatBuffers bytes of a tree, written by the oracle (its own
    layout) or by the reference's FlatBufferBuilder (impl='ref'; needs `sites` for the tree's reference sequence)."""
    n = e.num_nodes
    cap = 4096 + 16 * n + 16 * int(e.mut_off[n]) + 12 * int(e.miss_off[n]) + len(ref_seq)
    buf = (C.c_uint8 * cap)()
    es = e.as_struct()
    if impl == "ref":
        ss = sites.as_struct()
        got = ref().ref_api_tree_write(C.byref(es), C.byref(ss), buf, cap)
    else:
        r = np.ascontiguousarray(ref_seq, dtype=np.uint8)
        got = oracle().orc_api_tree_write(C.byref(es), _p(r, u8p), len(r), buf, cap)
    assert got > 0, got
    return bytes(buf[:got])


def api_tree_read(data: bytes, impl: str = "oracle"):
    """api_tree_and_tree_info_to_phylo_tree (core/api.cpp:127-186) -> (Emat, ref_seq).  impl='ref' runs the reference's reader (and its
    fix_up_missations); impl='oracle' the restatement, which raises ValueError(code) for buffers it does not take."""
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    counts = np.zeros(5, np.int32)
    if impl == "ref":
        lib = ref()
        nul = [None] * 15
        lib.ref_api_tree_read(buf, _p(counts, i32p), *nul, None)
        arrs = _alloc_emat_arrays(counts)
        view = OrcApiTreeView()
        if oracle().orc_api_tree_parse(buf, len(data), C.byref(view)) != 0:
            raise ValueError(-1)
        ref_seq = np.zeros(max(view.num_sites, 1), np.uint8)
        lib.ref_api_tree_read(buf, _p(counts, i32p), *_ptrs(arrs), _p(ref_seq, u8p))
        return _emat_from_arrays(counts, arrs), ref_seq[:view.num_sites]
    lib = oracle()
    view = OrcApiTreeView()
    rc = lib.orc_api_tree_parse(buf, len(data), C.byref(view))
    if rc != 0:
        raise ValueError(rc)
    nul = [None] * 15
    rc = lib.orc_api_tree_to_emat(C.byref(view), _p(counts, i32p), *nul)
    if rc != 0:
        raise ValueError(rc)
    arrs = _alloc_emat_arrays(counts)
    rc = lib.orc_api_tree_to_emat(C.byref(view), _p(counts, i32p), *_ptrs(arrs))
    if rc != 0:
        raise ValueError(rc)
    ref_seq = np.ctypeslib.as_array(C.cast(view.ref_seq, u8p), shape=(max(view.num_sites, 1),))[:view.num_sites].copy()
    return _emat_from_arrays(counts, arrs), ref_seq


def ref_maple_read(text: bytes):
    """read_maple (core/io.cpp:98-254) of the compiled reference -> the dict delphy_b200.maple_parse returns, or None where it throws."""
    lib = ref()
    counts = np.zeros(6, np.int64)
    i64p = C.POINTER(C.c_int64)
    nul = [None] * 12
    if lib.ref_maple_read(text, len(text), _p(counts, i64p), *nul) != 0:
        return None
    L, n, D, I, B, W = (int(c) for c in counts)
    ref_seq = np.zeros(max(L, 1), np.uint8); t_min = np.zeros(max(n, 1)); t_max = np.zeros(max(n, 1))
    name_off = np.zeros(n + 1, np.int64); names = C.create_string_buffer(max(B, 1))
    d_off = np.zeros(n + 1, np.int32); d_site = np.zeros(max(D, 1), np.int32); d_from = np.zeros(max(D, 1), np.uint8); d_to = np.zeros(max(D, 1), np.uint8)
    m_off = np.zeros(n + 1, np.int32); m_s = np.zeros(max(I, 1), np.int32); m_e = np.zeros(max(I, 1), np.int32)
    rc = lib.ref_maple_read(text, len(text), _p(counts, i64p), _p(ref_seq, u8p), _p(t_min, f64p), _p(t_max, f64p), _p(name_off, i64p), names,
                            _p(d_off, i32p), _p(d_site, i32p), _p(d_from, u8p), _p(d_to, u8p), _p(m_off, i32p), _p(m_s, i32p), _p(m_e, i32p))
    assert rc == 0
    raw = names.raw[:B]
    return dict(num_warnings=W, ref=ref_seq[:L], t_min=t_min[:n], t_max=t_max[:n], names=[raw[name_off[k]:name_off[k + 1]] for k in range(n)],
                delta_off=d_off, delta_site=d_site[:D], delta_from=d_from[:D], delta_to=d_to[:D], miss_off=m_off, miss_start=m_s[:I], miss_end=m_e[:I])


# ------------------------------------------------------------------------------------------------
def _build(target: str):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True, capture_output=True)


_ORACLE = None
_REF = None


def oracle() -> C.CDLL:
    global _ORACLE
    if _ORACLE is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ORACLE_DIR, "emat_oracle.c")):
            _build("oracle")
        lib = C.CDLL(path)
        E, S = C.POINTER(OrcEmat), C.POINTER(OrcSites)
        R = C.POINTER(OrcRegion)
        lib.orc_state_frequencies_per_partition.argtypes = [S, i32p]
        lib.orc_cum_Q_l.argtypes = [S, f64p]
        lib.orc_lambda_for_sequence.argtypes = [S]; lib.orc_lambda_for_sequence.restype = C.c_double
        lib.orc_lambda_i.argtypes = [E, S, f64p, f64p]
        lib.orc_log_root_prior.argtypes = [E, S, i32p]; lib.orc_log_root_prior.restype = C.c_double
        lib.orc_log_G_below_root.argtypes = [E, S, f64p]; lib.orc_log_G_below_root.restype = C.c_double
        lib.orc_branch_log_G.argtypes = [E, S, C.c_int32, C.c_double]; lib.orc_branch_log_G.restype = C.c_double
        lib.orc_path_log_G.argtypes = [E, S, C.c_int32, C.c_int32, f64p, i32p]; lib.orc_path_log_G.restype = C.c_double
        lib.orc_num_sites_missing_at_every_node.argtypes = [E, i32p]
        lib.orc_num_muts.argtypes = [E]; lib.orc_num_muts.restype = C.c_int32
        lib.orc_num_muts_ab.argtypes = [E, i32p]
        lib.orc_num_muts_beta_ab.argtypes = [E, S, i32p]
        lib.orc_num_muts_l.argtypes = [E, C.c_int32, i32p]
        lib.orc_num_muts_l_ab.argtypes = [E, C.c_int32, i32p]
        lib.orc_T.argtypes = [E]; lib.orc_T.restype = C.c_double
        lib.orc_T_l_a.argtypes = [E, S, f64p]
        lib.orc_Ttwiddle_l.argtypes = [E, S, f64p]
        lib.orc_Ttwiddle_beta_a.argtypes = [E, S, f64p]
        lib.orc_missing_sites_at.argtypes = [E, C.c_int32, i32p, i32p, C.c_int32]; lib.orc_missing_sites_at.restype = C.c_int32
        lib.orc_site_state_at_node.argtypes = [E, S, C.c_int32, C.c_int32]; lib.orc_site_state_at_node.restype = C.c_uint8
        lib.orc_spr_study_build.argtypes = [E, C.c_int32, C.c_int32, C.c_double, i32p, i32p, C.c_int32, C.c_int32, C.c_int32,
                                            i32p, u8p, u8p, C.c_int32, C.c_int32, C.c_int32, R, C.c_int32]
        lib.orc_spr_study_build.restype = C.c_int32
        lib.orc_spr_study_weights.argtypes = [E, C.c_int32, C.c_int32, R, C.c_int32, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.POINTER(OrcStudySummary)]
        lib.orc_spr_pick_nexus_region.argtypes = [R, C.c_int32, C.c_double]; lib.orc_spr_pick_nexus_region.restype = C.c_int32
        lib.orc_spr_find_region.argtypes = [R, C.c_int32, C.c_int32, C.c_double]; lib.orc_spr_find_region.restype = C.c_int32
        lib.orc_spr_log_alpha_in_region.argtypes = [E, R, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                                    C.c_double, C.c_double, C.c_double]
        lib.orc_spr_log_alpha_in_region.restype = C.c_double
        lib.orc_spr_study_from_attached.argtypes = [E, S, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, f64p,
                                                    R, C.c_int32, C.POINTER(OrcStudySummary)]
        lib.orc_spr_study_from_attached.restype = C.c_int32
        lib.orc_gamma_q_export.argtypes = [C.c_double, C.c_double]; lib.orc_gamma_q_export.restype = C.c_double
        lib.orc_api_tree_parse.argtypes = [C.c_void_p, C.c_int64, C.POINTER(OrcApiTreeView)]; lib.orc_api_tree_parse.restype = C.c_int32
        lib.orc_api_tree_to_emat.argtypes = [C.POINTER(OrcApiTreeView), i32p] + _EMAT_OUT_ARGS; lib.orc_api_tree_to_emat.restype = C.c_int32
        lib.orc_api_tree_write.argtypes = [E, u8p, C.c_int32, C.c_void_p, C.c_int64]; lib.orc_api_tree_write.restype = C.c_int64
        _ORACLE = lib
    return _ORACLE


def ref_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libdelphy_ref.so"))


def ref() -> C.CDLL:
    """The reference's own code (oracle/_ref/libdelphy_ref.so).  Raises if it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libdelphy_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = C.CDLL(path)
        E, S = C.POINTER(OrcEmat), C.POINTER(OrcSites)
        R = C.POINTER(OrcRegion)
        lib.ref_assert_integrity.argtypes = [E, S]
        lib.ref_state_frequencies_per_partition.argtypes = [S, i32p]
        lib.ref_cum_Q_l.argtypes = [S, f64p]
        lib.ref_lambda_i.argtypes = [E, S, f64p, f64p]
        lib.ref_log_root_prior.argtypes = [E, S]; lib.ref_log_root_prior.restype = C.c_double
        lib.ref_log_G_below_root.argtypes = [E, S]; lib.ref_log_G_below_root.restype = C.c_double
        lib.ref_path_log_G.argtypes = [E, S, C.c_int32, C.c_int32]; lib.ref_path_log_G.restype = C.c_double
        lib.ref_num_sites_missing_at_every_node.argtypes = [E, S, i32p]
        lib.ref_num_muts.argtypes = [E, S]; lib.ref_num_muts.restype = C.c_int32
        lib.ref_num_muts_ab.argtypes = [E, S, i32p]
        lib.ref_num_muts_beta_ab.argtypes = [E, S, i32p]
        lib.ref_num_muts_l.argtypes = [E, S, i32p]
        lib.ref_num_muts_l_ab.argtypes = [E, S, i32p]
        lib.ref_T.argtypes = [E, S]; lib.ref_T.restype = C.c_double
        lib.ref_T_l_a.argtypes = [E, S, f64p]
        lib.ref_Ttwiddle_l.argtypes = [E, S, f64p]
        lib.ref_Ttwiddle_beta_a.argtypes = [E, S, f64p]
        lib.ref_missing_sites_at.argtypes = [E, S, C.c_int32, i32p, i32p, C.c_int32]; lib.ref_missing_sites_at.restype = C.c_int32
        lib.ref_spr_study_build.argtypes = [E, S, C.c_int32, C.c_double, i32p, i32p, C.c_int32, C.c_int32, C.c_int32,
                                            i32p, u8p, u8p, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_double, C.c_double, C.c_double,
                                            R, C.c_int32, C.POINTER(OrcStudySummary)]
        lib.ref_spr_study_build.restype = C.c_int32
        lib.ref_spr_study_from_attached.argtypes = [E, S, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, f64p,
                                                    R, C.c_int32, C.POINTER(OrcStudySummary)]
        lib.ref_spr_study_from_attached.restype = C.c_int32
        lib.ref_bench_log_G.argtypes = [E, S, C.c_int32, C.c_int32, f64p]; lib.ref_bench_log_G.restype = C.c_double
        lib.ref_bench_spr.argtypes = [E, S, i32p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_int64)]
        lib.ref_bench_spr.restype = C.c_double
        lib.ref_api_tree_write.argtypes = [E, S, C.c_void_p, C.c_int64]; lib.ref_api_tree_write.restype = C.c_int64
        lib.ref_api_tree_read.argtypes = [C.c_void_p, i32p] + _EMAT_OUT_ARGS + [u8p]; lib.ref_api_tree_read.restype = C.c_int32
        i64p = C.POINTER(C.c_int64)
        lib.ref_maple_read.argtypes = [C.c_char_p, C.c_int64, i64p, u8p, f64p, f64p, i64p, C.c_char_p, i32p, i32p, u8p, u8p, i32p, i32p, i32p]
        lib.ref_maple_read.restype = C.c_int32
        _REF = lib
    return _REF


# ---- convenience wrappers (numpy in / numpy out) ----------------------------------------------------
class Oracle:
    """High-level numpy API over either liboracle.so (impl='oracle') or libdelphy_ref.so (impl='ref')."""

    def __init__(self, impl: str = "oracle"):
        self.impl = impl
        self.lib = oracle() if impl == "oracle" else ref()

    # -- phylo_tree_calc
    def state_frequencies(self, s: Sites) -> np.ndarray:
        out = np.zeros((s.num_partitions, 4), np.int32)
        fn = self.lib.orc_state_frequencies_per_partition if self.impl == "oracle" else self.lib.ref_state_frequencies_per_partition
        fn(C.byref(s.as_struct()), _p(out, i32p))
        return out

    def cum_Q_l(self, s: Sites) -> np.ndarray:
        out = np.zeros(s.num_sites + 1, np.float64)
        fn = self.lib.orc_cum_Q_l if self.impl == "oracle" else self.lib.ref_cum_Q_l
        fn(C.byref(s.as_struct()), _p(out, f64p))
        return out

    def lambda_i(self, e: Emat, s: Sites, cumQ: np.ndarray | None = None) -> np.ndarray:
        if cumQ is None:
            cumQ = self.cum_Q_l(s)
        out = np.zeros(e.num_nodes, np.float64)
        fn = self.lib.orc_lambda_i if self.impl == "oracle" else self.lib.ref_lambda_i
        fn(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(cumQ, f64p), _p(out, f64p))
        return out

    def log_root_prior(self, e: Emat, s: Sites) -> float:
        if self.impl == "oracle":
            fr = self.state_frequencies(s)
            return float(self.lib.orc_log_root_prior(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(fr, i32p)))
        return float(self.lib.ref_log_root_prior(C.byref(e.as_struct()), C.byref(s.as_struct())))

    def log_G_below_root(self, e: Emat, s: Sites, lambda_i: np.ndarray | None = None) -> float:
        if self.impl == "oracle":
            if lambda_i is None:
                lambda_i = self.lambda_i(e, s)
            return float(self.lib.orc_log_G_below_root(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(lambda_i, f64p)))
        return float(self.lib.ref_log_G_below_root(C.byref(e.as_struct()), C.byref(s.as_struct())))

    def path_log_G(self, e: Emat, s: Sites, A: int, B: int) -> float:
        if self.impl == "oracle":
            lam = self.lambda_i(e, s); fr = self.state_frequencies(s)
            return float(self.lib.orc_path_log_G(C.byref(e.as_struct()), C.byref(s.as_struct()), A, B, _p(lam, f64p), _p(fr, i32p)))
        return float(self.lib.ref_path_log_G(C.byref(e.as_struct()), C.byref(s.as_struct()), A, B))

    def nsmn(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros(e.num_nodes, np.int32)
        if self.impl == "oracle":
            self.lib.orc_num_sites_missing_at_every_node(C.byref(e.as_struct()), _p(out, i32p))
        else:
            self.lib.ref_num_sites_missing_at_every_node(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, i32p))
        return out

    def num_muts(self, e: Emat, s: Sites) -> int:
        if self.impl == "oracle":
            return int(self.lib.orc_num_muts(C.byref(e.as_struct())))
        return int(self.lib.ref_num_muts(C.byref(e.as_struct()), C.byref(s.as_struct())))

    def num_muts_ab(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros((4, 4), np.int32)
        if self.impl == "oracle":
            self.lib.orc_num_muts_ab(C.byref(e.as_struct()), _p(out, i32p))
        else:
            self.lib.ref_num_muts_ab(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, i32p))
        return out

    def num_muts_beta_ab(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros((s.num_partitions, 4, 4), np.int32)
        fn = self.lib.orc_num_muts_beta_ab if self.impl == "oracle" else self.lib.ref_num_muts_beta_ab
        fn(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, i32p))
        return out

    def num_muts_l(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros(s.num_sites, np.int32)
        if self.impl == "oracle":
            self.lib.orc_num_muts_l(C.byref(e.as_struct()), s.num_sites, _p(out, i32p))
        else:
            self.lib.ref_num_muts_l(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, i32p))
        return out

    def num_muts_l_ab(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros((s.num_sites, 4, 4), np.int32)
        if self.impl == "oracle":
            self.lib.orc_num_muts_l_ab(C.byref(e.as_struct()), s.num_sites, _p(out, i32p))
        else:
            self.lib.ref_num_muts_l_ab(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, i32p))
        return out

    def T(self, e: Emat, s: Sites) -> float:
        if self.impl == "oracle":
            return float(self.lib.orc_T(C.byref(e.as_struct())))
        return float(self.lib.ref_T(C.byref(e.as_struct()), C.byref(s.as_struct())))

    def T_l_a(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros((s.num_sites, 4), np.float64)
        fn = self.lib.orc_T_l_a if self.impl == "oracle" else self.lib.ref_T_l_a
        fn(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, f64p))
        return out

    def Ttwiddle_l(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros(s.num_sites, np.float64)
        fn = self.lib.orc_Ttwiddle_l if self.impl == "oracle" else self.lib.ref_Ttwiddle_l
        fn(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, f64p))
        return out

    def Ttwiddle_beta_a(self, e: Emat, s: Sites) -> np.ndarray:
        out = np.zeros((s.num_partitions, 4), np.float64)
        fn = self.lib.orc_Ttwiddle_beta_a if self.impl == "oracle" else self.lib.ref_Ttwiddle_beta_a
        fn(C.byref(e.as_struct()), C.byref(s.as_struct()), _p(out, f64p))
        return out

    def missing_sites_at(self, e: Emat, s: Sites, node: int):
        cap = int(e.miss_off[-1]) + 1
        st = np.zeros(cap, np.int32); en = np.zeros(cap, np.int32)
        if self.impl == "oracle":
            n = self.lib.orc_missing_sites_at(C.byref(e.as_struct()), node, _p(st, i32p), _p(en, i32p), cap)
        else:
            n = self.lib.ref_missing_sites_at(C.byref(e.as_struct()), C.byref(s.as_struct()), node, _p(st, i32p), _p(en, i32p), cap)
        assert n >= 0
        return st[:n].copy(), en[:n].copy()

    # -- spr_study
    def spr_study(self, e: Emat, s: Sites, X: int, t_X: float, missing, start_branch: int, start_mut_idx: int,
                  init_deltas, max_muts_from_start: int = 2**31 - 1, can_change_root: bool = True,
                  weights=None):
        """missing = (starts, ends); init_deltas = [(site, from, to)]; weights = None | (lambda_X, f, t_max_tip).
        Returns (regions structured array, summary | None)."""
        ms = np.ascontiguousarray(missing[0], np.int32); me = np.ascontiguousarray(missing[1], np.int32)
        isite = np.ascontiguousarray([d[0] for d in init_deltas], np.int32)
        ifrom = np.ascontiguousarray([d[1] for d in init_deltas], np.uint8)
        ito = np.ascontiguousarray([d[2] for d in init_deltas], np.uint8)
        cap = e.num_nodes + int(e.mut_off[-1]) + 8
        out = np.zeros(cap, REGION_DTYPE)
        summ = OrcStudySummary()
        R = C.POINTER(OrcRegion)
        es, ss = e.as_struct(), s.as_struct()
        if self.impl == "oracle":
            n = self.lib.orc_spr_study_build(C.byref(es), s.num_sites, X, t_X, _p(ms, i32p), _p(me, i32p), len(ms),
                                             start_branch, start_mut_idx, _p(isite, i32p), _p(ifrom, u8p), _p(ito, u8p),
                                             len(isite), max_muts_from_start, int(can_change_root), _p(out, R), cap)
            assert n >= 0
            if weights is not None and n > 0:
                nmiss = int((me - ms).sum())
                self.lib.orc_spr_study_weights(C.byref(es), s.num_sites, nmiss, _p(out, R), n, weights[0], weights[1], t_X,
                                               weights[2], C.byref(summ))
        else:
            w = weights if weights is not None else (0.0, 0.0, 0.0)
            n = self.lib.ref_spr_study_build(C.byref(es), C.byref(ss), X, t_X, _p(ms, i32p), _p(me, i32p), len(ms),
                                             start_branch, start_mut_idx, _p(isite, i32p), _p(ifrom, u8p), _p(ito, u8p),
                                             len(isite), max_muts_from_start, int(can_change_root),
                                             int(weights is not None), w[0], w[1], w[2], _p(out, R), cap, C.byref(summ))
            assert n >= 0
        return out[:n].copy(), (summ if weights is not None else None)

    def spr_study_from_attached(self, e: Emat, s: Sites, X: int, lambda_i: np.ndarray, max_muts_from_start=2**31 - 1,
                                can_change_root=True, annealing_factor=0.8, t_max_tip=None):
        if t_max_tip is None:
            t_max_tip = float(e.t.max())
        cap = e.num_nodes + int(e.mut_off[-1]) + 8
        out = np.zeros(cap, REGION_DTYPE)
        summ = OrcStudySummary()
        fn = self.lib.orc_spr_study_from_attached if self.impl == "oracle" else self.lib.ref_spr_study_from_attached
        n = fn(C.byref(e.as_struct()), C.byref(s.as_struct()), X, max_muts_from_start, int(can_change_root),
               annealing_factor, t_max_tip, _p(lambda_i, f64p), _p(out, C.POINTER(OrcRegion)), cap, C.byref(summ))
        assert n >= 0
        return out[:n].copy(), summ
