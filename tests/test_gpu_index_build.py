"""K5: the CUDA index builder yields the same artefact as the host builder and the oracle's
independent builder — same unitigs / colours / exts (text dump), same flat arrays (nb_index_compare), and reads aligned
through a GPU-built index give the oracle's counts.  Through the C ABI (nb_index_build_gpu*)."""
import json
import os
import random

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth
from tests.test_gpu_parity import compare, make

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "ref")
OCFG = dict(score_percent=0.1, score_threshold=50, num_mismatches=0, discard_multiple_matches=False, require_valid_pair=False,
            discard_multi_hits=0, max_hits_to_report=5, intersect_level=0, strand_filter="none", trim_target_length=15, trim_strictness=0.5)


def _parity(seqs, threads=4):
    names = ["s%d" % i for i in range(len(seqs))]
    o = orc.Oracle(OCFG, orc.Reference(0, ["sequence_name", "sequence"], [names, seqs], 0, 1))
    host = nb.Index.from_sequences(seqs, threads)
    dev = nb.Index.from_sequences(seqs, threads, device=0)
    assert dev.stats() == host.stats()
    so = o.index_stats()
    for k in so:
        assert so[k] == dev.stats()[k], k
    assert dev.dump() == o.index_dump()
    assert dev.compare(host) == 0 and host.compare(dev) == 0
    return host, dev


@pytest.mark.parametrize("name", ["basic.json", "basic-rev.json", "mismatch.json", "strandedness.json"])
def test_gpu_index_matches_host_and_oracle_on_reference_fixtures(name):
    _, ref = orc.get_reference_library(os.path.join(G, "libraries", name), "none")
    _parity(ref.columns[ref.sequence_idx])


def test_gpu_index_forks_shared_segments_short_and_non_acgt():
    rnd = random.Random(11)
    rs = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    core, core2 = rs(60), rs(31)
    seqs = [rs(29), "", rs(30), core + rs(40), rs(35) + core + rs(20), core, rs(10) + core2 + rs(10), core2 + rs(3), rs(200) + core2,
            "ACGTNNNNACGTRYKM" * 5 + rs(40), rs(500), "acgtacgtacgtacgtacgtacgtacgtacgtacgtacgt"[:33] + rs(20), core + rs(40)]
    _parity(seqs)


def test_gpu_index_pure_cycles_fall_back_to_the_host_rule():
    rnd = random.Random(5)
    rs = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    seqs = ["A" * 50, "AC" * 40, "CA" * 40, "ACG" * 30, rs(100), (rs(45) * 3)[:120]]
    _parity(seqs, threads=2)


def test_gpu_index_compare_detects_differences():
    rnd = random.Random(6)
    rs = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    a = [rs(100), rs(80)]
    b = [a[0], a[1][:-1] + ("A" if a[1][-1] != "A" else "C")]
    assert nb.Index.from_sequences(a, 1, device=0).compare(nb.Index.from_sequences(b, 1)) != 0


def test_gpu_index_no_device_is_an_error():
    with pytest.raises(nb.NbError) as e:
        nb.Index.from_sequences(["ACGT" * 20], 1, device=99)
    assert e.value.code == -6


def test_gpu_built_index_aligns_like_the_oracle():
    """C2-shaped family library, both group_on settings: build on the GPU, align, compare everything with the oracle."""
    L = synth.SynthLibrary(seed=1234, n_fam=200, n_all=5, group_on="")
    obj = L.to_json_obj()
    for group_on in ("", "family"):
        ocfg, oref, lib = make(obj, "unstranded", group_on)
        dev = nb.build_index(lib, 8, device=0)
        host = nb.build_index(lib, 8)
        assert dev.compare(host) == 0
        o = orc.Oracle(ocfg, oref)
        assert dev.dump() == o.index_dump()
        ctx = nb.Context(dev, lib)
        for mm in (0, 2):
            cfg = dict(ocfg, num_mismatches=mm)
            o.set_config(**cfg)
            r1, o1, r2, o2 = synth.pairs(L, 0, 20000, seed=31 + mm, paired=True)
            compare(ctx, o, cfg, r1, o1, r2, o2)


def test_gpu_index_big_components_and_save_load(tmp_path):
    Lbig = synth.SynthLibrary(seed=99, n_fam=2, n_all=90)
    rng = np.random.default_rng(17)
    for f in range(2):
        base = Lbig.seqs[int(Lbig.off[f * 90]):int(Lbig.off[f * 90 + 1])].copy()
        for a in range(90):
            s = base.copy()
            for pos in rng.integers(0, len(s), size=2):
                s[pos] = ord("ACGT"[(("ACGT".index(chr(s[pos]))) + 1 + int(rng.integers(3))) % 4])
            Lbig.seqs[int(Lbig.off[f * 90 + a]):int(Lbig.off[f * 90 + a + 1])] = s
    ocfg, oref, lib = make(Lbig.to_json_obj(), "unstranded", "family")
    dev = nb.build_index(lib, 4, device=0)
    assert dev.compare(nb.build_index(lib, 4)) == 0
    assert dev.dump() == orc.Oracle(ocfg, oref).index_dump()
    dev.save(tmp_path / "big.nbidx")
    assert nb.Index.load(tmp_path / "big.nbidx").compare(dev) == 0


def test_index_cache_miss_then_hit(tmp_path):
    """nb_index_build_cached (what the file drivers call): a miss builds on the GPU and leaves <key>.nbix behind, the next
    call takes that file; both are the host builder's artefact."""
    _, lib = nb.get_reference_library(os.path.join(G, "libraries", "basic.json"))
    cache = tmp_path / "cache"; cache.mkdir()
    a = nb.Index.build_cached(lib, cache, device=0, threads=2)
    key = nb.Index.cache_key(lib)
    assert sorted(x.name for x in cache.iterdir()) == [key + ".nbix"]
    stamp = (cache / (key + ".nbix")).stat().st_mtime_ns
    b = nb.Index.build_cached(lib, cache, device=0, threads=2)
    assert (cache / (key + ".nbix")).stat().st_mtime_ns == stamp          # not rebuilt
    host = nb.build_index(lib, 2)
    assert a.compare(host) == 0 and b.compare(host) == 0
    # a cache directory that cannot be written is not an error: the index is built and used
    c = nb.Index.build_cached(lib, tmp_path / "does" / "not" / "exist", device=0, threads=2)
    assert c.compare(host) == 0
