"""Loud-failure behaviour of the device path: every capacity limit and unsupported input returns an error code with a
message (the reference panics with a message; a silent wrong answer is never acceptable)."""
import json

import numpy as np
import pytest

import nimble_aligner_b200 as nb
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    L = synth.SynthLibrary(seed=3, n_fam=30, n_all=5)
    lib = nb.Library.from_text(json.dumps(L.to_json_obj()), "unstranded")
    ix = nb.build_index(lib, 4)
    r1, o1, r2, o2 = synth.pairs(L, 0, 20000)
    return L, lib, ix, (r1, o1, r2, o2)


def test_capacity_overflows_are_reported(small):
    L, lib, ix, (r1, o1, r2, o2) = small
    ctx = nb.Context(ix, lib, callset_slots=16)
    ctx.align_batch(r1, o1, r2, o2)
    with pytest.raises(nb.NbError) as e:
        ctx.counts()
    assert e.value.code == -7 and "callset_slots" in str(e.value)
    scope = (np.arange(20000) // 3).astype(np.uint32)
    ctx = nb.Context(ix, lib, agg_slots=16)
    ctx.align_batch(r1, o1, r2, o2, scope_id=scope)
    with pytest.raises(nb.NbError) as e:
        ctx.counts()
    assert e.value.code == -7 and "agg_slots" in str(e.value)


def test_arena_overflow_is_reported():
    Lbig = synth.SynthLibrary(seed=99, n_fam=1, n_all=100)
    base = Lbig.seqs[int(Lbig.off[0]):int(Lbig.off[1])].copy()
    for a in range(100):   # 100 identical alleles: every colour has 100 ids -> arena path
        Lbig.seqs[int(Lbig.off[a]):int(Lbig.off[a + 1])] = base
    lib = nb.Library.from_text(json.dumps(Lbig.to_json_obj()), "unstranded")
    ix = nb.build_index(lib, 2)
    r1, o1, r2, o2 = synth.pairs(Lbig, 0, 5000, seed=5)
    ok = nb.Context(ix, lib)
    reads, _ = ok.align_batch(r1, o1, r2, o2, want_reads=True)
    assert reads["ec_len"].max() == 100
    ctx = nb.Context(ix, lib, ec_arena_entries=1024)
    ctx.align_batch(r1, o1, r2, o2)
    with pytest.raises(nb.NbError) as e:
        ctx.counts()
    assert e.value.code == -7 and "ec_arena_entries" in str(e.value)


def test_unsupported_inputs_fail_loudly(small):
    L, lib, ix, (r1, o1, r2, o2) = small
    ctx = nb.Context(ix, lib)
    with pytest.raises(nb.NbError) as e:
        ctx.set_config(lib.config.copy(max_hits_to_report=64))
    assert e.value.code == -5
    with pytest.raises(nb.NbError) as e:   # > 1024 bases
        d, o = nb.pack_reads(["ACGT" * 300])
        ctx.align_batch(d, o)
    assert e.value.code == -5
    # duplicate sequence_name: the reference maps both to the first row; the device tables cannot -> refuse
    cfg = lib.config
    seqs = L.sequences()[:2]
    dup = nb.Library.from_columns(["sequence_name", "sequence"], [["X", "X"], seqs], 0, 0, 1, cfg)
    dix = nb.build_index(dup, 1)
    dctx = nb.Context(dix, dup)
    with pytest.raises(nb.NbError) as e:
        dctx.align_batch(r1, o1)
    assert e.value.code == -5 and "duplicate sequence_name" in str(e.value)
    # a library holding only a §rev row: unmap() of its base name panics in the reference ("Feature not found ...")
    only_rev = nb.Library.from_columns(["sequence_name", "sequence"], [["Y" + "§" + "rev"], [seqs[0]]], 0, 0, 1, cfg)
    rix = nb.build_index(only_rev, 1)
    rctx = nb.Context(rix, only_rev)
    d, o = nb.pack_reads([seqs[0][100:250]])
    rctx.align_batch(d, o)
    with pytest.raises(nb.NbError) as e:
        rctx.counts()
    assert e.value.code == -8 and "Feature not found" in str(e.value)


def test_state_machine_misuse_is_rejected(small):
    L, lib, ix, (r1, o1, r2, o2) = small
    ctx = nb.Context(ix, lib)
    ctx.align_batch(r1, o1, r2, o2)
    with pytest.raises(nb.NbError):   # scoped after whole-run without reset
        ctx.align_batch(r1, o1, r2, o2, scope_id=np.zeros(20000, dtype=np.uint32))
    ctx.counts()
    with pytest.raises(nb.NbError):   # more batches after finalize without reset
        ctx.align_batch(r1, o1, r2, o2)
    ctx.reset()
    ctx.align_batch(r1, o1, r2, o2)
    a = ctx.counts()["rows"]
    assert ctx.counts()["rows"] == a   # finalize is idempotent


def test_fastq_driver_input_errors(tmp_path, small):
    """process::fastq::process panics on malformed / unequal inputs (src/process/fastq.rs:20-24, src/parse/fastq.rs:30-33):
    here a negative status and no TSV.  Also multi-line records, CRLF, blank lines and a missing final newline parse."""
    L, lib, ix, _ = small
    libp = tmp_path / "lib.json"
    libp.write_text(json.dumps(L.to_json_obj()))
    seqs = L.sequences()
    rec = lambda i, s: "@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s))
    good = "".join(rec(i, seqs[i % len(seqs)][10:160]) for i in range(50))
    (tmp_path / "a.fastq").write_text(good)
    (tmp_path / "short.fastq").write_text("".join(rec(i, seqs[i % len(seqs)][10:160]) for i in range(49)))
    (tmp_path / "bad.fastq").write_text(good[:-40] + "garbage\n")
    (tmp_path / "noat.fastq").write_text(good.replace("@r7\n", "r7\n"))
    out = tmp_path / "o.tsv"
    for r2 in ("short.fastq", "bad.fastq", "noat.fastq"):
        with pytest.raises(nb.NbError) as e:
            nb.process_fastq([tmp_path / "a.fastq", tmp_path / r2], [libp], [out])
        assert e.value.code == -3 and not out.exists()
    with pytest.raises(nb.NbError):
        nb.process_fastq([tmp_path / "missing.fastq"], [libp], [out])
    # tolerant parsing: sequence split over two lines, CRLF, blank line between records, no newline at the very end
    s0, s1 = seqs[3][20:170], seqs[4][5:155]
    (tmp_path / "odd.fastq").write_text("@x\r\n%s\r\n%s\r\n+\r\n%s\r\n%s\r\n\n@y\n%s\n+y\n%s" % (s0[:70], s0[70:], "I" * 70, "I" * 80, s1, "I" * 150))
    (tmp_path / "plain.fastq").write_text(rec(0, s0) + rec(1, s1))
    nb.process_fastq([tmp_path / "odd.fastq"], [libp], [tmp_path / "odd.tsv"])
    nb.process_fastq([tmp_path / "plain.fastq"], [libp], [tmp_path / "plain.tsv"])
    assert (tmp_path / "odd.tsv").read_text() == (tmp_path / "plain.tsv").read_text() and len((tmp_path / "odd.tsv").read_text().splitlines()) >= 2
