"""CPU-only: dphy_maple_parse (delphy_b200/csrc/maple.cpp) against the reference's own reader (read_maple, core/io.cpp:98-254) compiled
in place -- same reference sequence, same surviving samples with the same names / date ranges / substitutions / merged missing runs,
same number of warnings, same refusals -- on hand-written edge cases, on the synthetic alignments of the bench, and on randomly
damaged files; plus a committed fixture (text + what the reference read from it: tests/golden/maple_fixture.*) for boxes
without the compiled reference."""
import json
import os

import numpy as np
import pytest

import delphy_b200 as db
import oracle_lib as ol
from delphy_b200.maple import write_maple
from helpers import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARRAYS = ("ref", "t_min", "t_max", "delta_off", "delta_site", "delta_from", "delta_to", "miss_off", "miss_start", "miss_end")

EDGE_CASES = b""">reference id is ignored
ACGTNACGTRYACGT
ACGTTTTTACGT

>plain|2021-03-04
c 1
n 5 3
T\t20
t 21
u 2
-\t9
>impossible_date|2021-02-30
c 2
>leap_day|2020-02-29
a 4
>not_a_leap_day|2021-02-29
a 4
>whole_month-2020-02
g 4
n 10
n 11 2
n 3 1
n 12 4
n 20 1
>whole_year|2019
c 1
>no_date_at_all
c 1
>bad_letter|2019-05-05
x 3
>range|2020-01-05/2020-02-07
a 2
>range_with_bad_end|2020-01-05/2020-02-30
a 2
>month_thirteen|2020-13
>same_as_reference|2021-01-01
a 1
>t_to_t_is_skipped|2021-01-01
t 4
T 8
c 4
>ambiguous_reference_site|2021-01-01
c 5
>site_past_the_end|2021-01-01
c 100
>run_before_the_start|2021-01-01
n 0 2
>run_to_the_very_end|2021-01-01
n 27 1
n 1 27
>run_past_the_end|2021-01-01
n 27 2
>junk_after_a_site|2021-01-01
c 6 junk
>junk_run_length|2021-01-01
n 7 zz
>negative_length|2021-01-01
n 7 -2
>plus_sign|2021-01-01
c +6
n +7 +2
>huge_number|2021-01-01
c 99999999999999999999
>lower_case_and_iupac|2021-01-01
r 3 2
y 8
? 1
. 2 1
>  blanks around the name |2020-06-15  \r
G 2\r
\r
   \t
>only_a_blank_line|2021-01-01
 
>letter_only|2021-01-01
c
>run_letter_only|2021-01-01
n
>no_lines|2021-01-01
>last_one_without_newline|2021-01-01
c 1"""


def _ours(text):
    try:
        return db.maple_parse(text)
    except db.DphyError:
        return None


def _same(a, b):
    assert (a is None) == (b is None)
    if a is None:
        return
    assert a["names"] == b["names"]
    assert a["num_warnings"] == b["num_warnings"]
    for k in ARRAYS:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and np.array_equal(x, y), k


def test_committed_fixture():
    text = open(os.path.join(GOLDEN, "maple_fixture.maple"), "rb").read()
    want = json.load(open(os.path.join(GOLDEN, "maple_fixture.json")))
    got = db.maple_parse(text)
    assert [n.decode() for n in got["names"]] == want["names"] and got["num_warnings"] == want["num_warnings"]
    for k in ARRAYS:
        assert np.array_equal(np.asarray(got[k], np.float64), np.asarray(want[k], np.float64)), k
    assert len(got["names"]) >= 12          # the fixture is the edge-case file: most of its samples are the ones the reference keeps


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref/libdelphy_ref.so not built")
def test_edge_cases_against_the_compiled_reference():
    want = ol.ref_maple_read(EDGE_CASES)
    _same(_ours(EDGE_CASES), want)
    kept = [n.decode() for n in want["names"]]
    for name in ("plain|2021-03-04", "leap_day|2020-02-29", "whole_year|2019", "range|2020-01-05/2020-02-07", "t_to_t_is_skipped|2021-01-01",
                 "run_to_the_very_end|2021-01-01", "plus_sign|2021-01-01", "lower_case_and_iupac|2021-01-01", "no_lines|2021-01-01",
                 "last_one_without_newline|2021-01-01"):
        assert name in kept, name
    for name in ("impossible_date|2021-02-30", "not_a_leap_day|2021-02-29", "no_date_at_all", "same_as_reference|2021-01-01",
                 "ambiguous_reference_site|2021-01-01", "run_past_the_end|2021-01-01", "junk_run_length|2021-01-01", "letter_only|2021-01-01"):
        assert name not in kept, name
    # with and without a final newline, CRLF throughout
    _same(_ours(EDGE_CASES + b"\n"), ol.ref_maple_read(EDGE_CASES + b"\n"))
    crlf = EDGE_CASES.replace(b"\r", b"").replace(b"\n", b"\r\n")
    _same(_ours(crlf), ol.ref_maple_read(crlf))


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref/libdelphy_ref.so not built")
@pytest.mark.parametrize("text", [b"", b"\n", b"ACGT\n", b">ref", b">ref\n", b">ref\nACGT", b">ref\nACGT\n", b">ref\nACXT\n>a|2020-01-01\n",
                                  b">ref\n\n\nAC\nGT\n\n", b">ref\nAC GT\t\r\n>a|2020-01-01\nc 1\n", b">ref\n>a|2020-01-01\nc 1\n",
                                  b">ref\nACGT\n>a|2020-01-01\n\n\n>b|2020-01-02\n"])
def test_headers_and_refusals_against_the_compiled_reference(text):
    _same(_ours(text), ol.ref_maple_read(text))


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref/libdelphy_ref.so not built")
@pytest.mark.parametrize("cfg,ov", [(0, dict(num_tips=300)), (5, dict(num_tips=150, num_sites=40000)), (2, dict(num_tips=400))])
def test_synthetic_alignments_against_the_compiled_reference(tmp_path, cfg, ov):
    emat, sites, info = synth(cfg, **ov)
    path = str(tmp_path / "a.maple")
    n = write_maple(emat, sites, path, info["t_max_tip"])
    text = open(path, "rb").read()
    got, want = _ours(text), ol.ref_maple_read(text)
    _same(got, want)
    assert len(got["names"]) == n and got["num_warnings"] == 0 and np.array_equal(got["ref"], sites.ref)
    assert int(got["miss_off"][-1]) > 0 and int(got["delta_off"][-1]) > 0


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref/libdelphy_ref.so not built")
def test_randomly_damaged_files_against_the_compiled_reference(tmp_path):
    emat, sites, info = synth(0, num_tips=60, num_sites=400, muts_per_tip=4.0)
    path = str(tmp_path / "a.maple")
    write_maple(emat, sites, path, info["t_max_tip"])
    base = open(path, "rb").read() + EDGE_CASES[EDGE_CASES.index(b">plain"):]
    rng = np.random.default_rng(11)
    noise = b"acgtnACGTN-?.xyz0123456789 \t\r\n>|-/+"
    for _ in range(400):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(0, len(b)))
            op = int(rng.integers(0, 3))
            if op == 0:
                b[at] = noise[int(rng.integers(0, len(noise)))]
            elif op == 1:
                b[at:at] = bytes(noise[int(rng.integers(0, len(noise)))] for _ in range(int(rng.integers(1, 4))))
            else:
                del b[at:at + int(rng.integers(1, 5))]
        text = bytes(b)
        _same(_ours(text), ol.ref_maple_read(text))
