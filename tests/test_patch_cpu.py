"""The host side of snaq_patch_network on the CPU tier (tests/hostsim): after an edit of the network, a quartet none of
whose taxa is reported as touched must keep its program, its length and its header on the re-rooted new tables -- that
is what lets the device keep its staged program, sort key and length (snaq.jl_b200/csrc/snaq_tables.cpp:
snaq_tables_touched; src/snaq_optimization.jl:1154-1198 is the reference's in-place edit of the candidate)."""
import ctypes as C

import numpy as np
import pytest

import snaq_jl_b200 as S
from snaq_jl_b200 import simulate
from tests import hostsim_py as hs


STATS = []


def progs(flat_root, flat, n, qt):
    nq = len(qt)
    words = np.zeros((nq, 8), np.uint16)
    ln = np.zeros(nq, np.int32)
    hdr = np.zeros(nq, np.uint8)
    rc = hs.lib().hostsim_progs_rooted(flat_root.ptr(), flat.ptr(), n, C.c_int64(nq), hs._p(np.ascontiguousarray(qt, dtype=np.uint16)),
                                       hs._p(words), hs._p(ln), hs._p(hdr))
    assert rc == 0, hs.lib().hostsim_last_error()
    return words, ln, hdr


@pytest.mark.parametrize("seed", range(12))
def test_quartets_without_a_touched_taxon_keep_their_programs(seed):
    n, h = 12 + seed % 9, seed % 4
    net = None
    for attempt in range(20):
        try:
            net = simulate.random_level1_network(n, h, 4100 + seed + 1000 * attempt)
            break
        except S.SnaqError:
            continue
    taxa = [f"t{i + 1}" for i in range(n)]
    qt = simulate.all_quartets(n)
    rng = np.random.default_rng(seed)
    partial = changed = 0
    for move in range(6):
        old = net.flatten(taxa)
        if simulate.nni_move(net, rng) is None:
            break
        new = net.flatten(taxa)
        touched = np.zeros(n, np.uint8)
        rc = hs.lib().hostsim_touched(old.ptr(), new.ptr(), n, hs._p(touched))
        assert rc >= 0, hs.lib().hostsim_last_error()
        if rc == 0:
            continue                                     # full classification: nothing to check
        partial += 1
        w0, l0, h0 = progs(old, old, n, qt)
        w1, l1, h1 = progs(old, new, n, qt)
        dirty = touched[qt].any(axis=1)
        same = (w0 == w1).all(axis=1) & (l0 == l1) & (h0 == h1)
        assert same[~dirty].all(), (seed, move, int((~same & ~dirty).sum()))
        changed += int((~same).sum())
        assert (~same).sum() <= dirty.sum()
    STATS.append((partial, changed))


def test_the_sweep_above_did_exercise_partial_patches():
    assert sum(p for p, _ in STATS) >= 10 and sum(c for _, c in STATS) > 0
