"""MAPLE text for a synthetic EMAT, so that the reference's own CLI (tools/delphy.cpp, --v0-in-maple) can run on the synthetic
alignments of BASELINE.json (config 1 "also written as MAPLE for the stock CLI").

Format as parsed by the reference (core/io.cpp:98-255): `>ref` + reference sequence, then per tip `>name|YYYY-MM-DD`
followed by one line per difference from the reference (`<letter> <1-based site>`) or gap (`n <1-based start> <length>`).
Tip dates come from the sequence id (core/sequence_utils.cpp:98-160); the reference measures time in days from 2020-01-01
(core/dates.cpp:12-21), the synthetic generator in years before the latest tip, which is pinned to 2021-01-01 here.
"""
import datetime

import numpy as np

_LETTERS = "ACGT"
_LATEST_TIP = datetime.date(2021, 1, 1)


def tip_sequences(emat, sites):
    """Yields (tip node index, {site: state} differences from the reference sequence, [(start, end)] missing intervals)."""
    ref = sites.ref
    N = emat.num_nodes
    # iterative DFS carrying the difference map and the missing intervals accumulated from the root
    stack = [(emat.root, {}, [])]
    while stack:
        v, diffs, missing = stack.pop()
        diffs = dict(diffs)
        for i in range(int(emat.mut_off[v]), int(emat.mut_off[v + 1])):
            l, to = int(emat.mut_site[i]), int(emat.mut_to[i])
            if to == int(ref[l]):
                diffs.pop(l, None)
            else:
                diffs[l] = to
        m0, m1 = int(emat.miss_off[v]), int(emat.miss_off[v + 1])
        if m1 > m0:
            missing = missing + [(int(emat.miss_start[i]), int(emat.miss_end[i])) for i in range(m0, m1)]
        if emat.child0[v] < 0:
            yield v, diffs, sorted(missing)
        else:
            stack.append((int(emat.child0[v]), diffs, missing))
            stack.append((int(emat.child1[v]), diffs, missing))
    assert N >= 1


def write_maple(emat, sites, path, t_max_tip=None):
    """Writes the alignment implied by `emat` as a MAPLE file; returns the number of tips written."""
    if t_max_tip is None:
        tips = np.nonzero(emat.child0 < 0)[0]
        t_max_tip = float(emat.t[tips].max())
    n = 0
    with open(path, "w") as f:
        f.write(">ref\n")
        f.write("".join(_LETTERS[int(c)] for c in sites.ref))
        f.write("\n")
        for v, diffs, missing in tip_sequences(emat, sites):
            days = int(round((float(emat.t[v]) - t_max_tip) * 365.0))
            date = _LATEST_TIP + datetime.timedelta(days=days)
            f.write(f">synth_tip_{v}|{date.isoformat()}\n")
            entries = []
            for s, e in missing:
                entries.append((s, f"n\t{s + 1}\t{e - s}\n"))
            for l, to in diffs.items():
                if any(s <= l < e for s, e in missing):
                    continue
                entries.append((l, f"{_LETTERS[to]}\t{l + 1}\n"))
            for _, line in sorted(entries):
                f.write(line)
            n += 1
    return n
