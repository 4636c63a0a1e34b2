"""Shared helpers for the parity tests: synthetic EMATs (product-side generator) -> oracle structs."""
from __future__ import annotations

import functools

import numpy as np

import delphy_b200 as db
from oracle_lib import Emat, Sites


def to_oracle(emat: db.HostEmat, sites: db.HostSites):
    e = Emat(emat.root, emat.parent, emat.child0, emat.child1, emat.t, emat.mut_off, emat.mut_site, emat.mut_from,
             emat.mut_to, emat.mut_t, emat.miss_off, emat.miss_start, emat.miss_end, emat.fs_off, emat.fs_site,
             emat.fs_from, emat.includes_run_root)
    s = Sites(sites.ref, sites.partition_for_site, sites.nu_l, sites.mu, sites.pi_a, sites.q_ab)
    return e, s


def from_oracle(e: Emat, s: Sites):
    emat = db.HostEmat(e.root, e.includes_run_root, parent=e.parent, child0=e.child0, child1=e.child1, t=e.t,
                       mut_off=e.mut_off, mut_site=e.mut_site, mut_from=e.mut_from, mut_to=e.mut_to, mut_t=e.mut_t,
                       miss_off=e.miss_off, miss_start=e.miss_start, miss_end=e.miss_end, fs_off=e.fs_off,
                       fs_site=e.fs_site, fs_from=e.fs_from)
    sites = db.HostSites(s.ref, s.partition_for_site, s.nu_l, s.mu, s.pi_a, s.q_ab)
    return emat, sites


@functools.lru_cache(maxsize=32)
def synth(config=0, **overrides):
    p = db.synth_params(config, **overrides)
    return db.synth_generate(p)


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    denom = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0
