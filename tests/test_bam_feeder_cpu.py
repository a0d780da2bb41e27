"""CPU test of the BAM feeder: the C++ BGZF/BAM reader + SortedBamReader/UMIReader logic behind
the C ABI must produce exactly the groups, clipped sequences, quals and 38 metadata fields that the independent Python
restatement (oracle/bam_ref.py) produces from the same file."""
import os

import pytest

import nimble_aligner_b200 as nb
import synth
from oracle import bam_ref
from tests.bamcases import make_bam


@pytest.mark.parametrize("no_mmap", [False, True])
@pytest.mark.parametrize("force_paired", [False, True])
def test_feeder_groups_match_python_restatement(tmp_path, monkeypatch, force_paired, no_mmap):
    if no_mmap:
        monkeypatch.setenv("NB_BAM_NO_MMAP", "1")      # the read() path kept for inputs that cannot be mapped (pipes)
    L = synth.SynthLibrary(seed=1234, n_fam=20, n_all=5)
    bam = make_bam(str(tmp_path / "t.bam"), L, n_groups=200)
    out = str(tmp_path / "groups.tsv")
    nb.bam_dump_groups(bam, out, force_bam_paired=force_paired, num_cores=3)
    got = [l.rstrip("\n").split("\t") for l in open(out, encoding="latin1")]
    groups = bam_ref.groups_of(bam_ref.read_bam(bam), force_paired)
    want = []
    for gi, g in enumerate(groups):
        for it in g:
            rev = it["f"][2] == "true"
            q_unrev = it["qual"][::-1] if rev else it["qual"]
            f = list(it["f"]); f[1] = it["qual"].hex()
            want.append([str(gi), it["seq"], q_unrev.hex()] + f)
    assert len(got) == len(want) and len(groups) > (20 if force_paired else 100)
    for a, b in zip(got, want):
        assert a == b
    # the quirks are really exercised
    flat = [it for g in groups for it in g]
    if not force_paired:
        assert any(it["f"][37] == "TRUE" for it in flat) and any(it["f"][4] == "true" for it in flat)
        assert any(len(it["seq"]) == 111 for it in flat)              # TSO clip
        assert any(it["f"][36] == "" for it in flat)                  # UB missing -> grouped by UR
    else:
        assert all(it["f"][37] == "" and it["f"][4] == "true" for it in flat)   # -p: no dummies, no SKIP_ALIGN aux


@pytest.mark.parametrize("force_paired", [False, True])
def test_parallel_grouping_equals_the_serial_readers(tmp_path, force_paired):
    """collect_groups restates SortedBamReader / UMIReader over whole UMI runs on several threads; the serial readers
    (kept in the library, NB_BAM_SERIAL_GROUPING=1) define the semantics.  Same dump on a BAM with thousands of groups."""
    L = synth.SynthLibrary(seed=99, n_fam=20, n_all=5)
    bam = make_bam(str(tmp_path / "big.bam"), L, n_groups=4000, seed=5)
    a, b = str(tmp_path / "par.tsv"), str(tmp_path / "ser.tsv")
    nb.bam_dump_groups(bam, a, force_bam_paired=force_paired, num_cores=8)
    os.environ["NB_BAM_SERIAL_GROUPING"] = "1"
    try:
        nb.bam_dump_groups(bam, b, force_bam_paired=force_paired, num_cores=8)
    finally:
        del os.environ["NB_BAM_SERIAL_GROUPING"]
    pa, se = open(a, "rb").read(), open(b, "rb").read()
    assert pa == se and pa.count(b"\n") > (500 if force_paired else 5000)


def test_feeder_fails_loudly_on_damaged_bam(tmp_path):
    """A BAM is outside input: truncated files, flipped bytes in the compressed stream and damaged record fields inside
    intact BGZF blocks must end in groups or an NbError (the reference panics inside htslib / on unwrap), never in a crash."""
    import random
    import struct
    import zlib
    L = synth.SynthLibrary(seed=1234, n_fam=20, n_all=5)
    base = open(make_bam(str(tmp_path / "t.bam"), L, n_groups=30), "rb").read()
    payload, i = b"", 0
    while i < len(base):                                   # BGZF blocks -> the BAM byte stream
        xlen = struct.unpack_from("<H", base, i + 10)[0]; bsize = struct.unpack_from("<H", base, i + 16)[0] + 1
        payload += zlib.decompress(base[i + 12 + xlen:i + bsize - 8], -15); i += bsize

    def bgzf(data):
        out = b""
        for k in range(0, max(len(data), 1), 20000):
            c = data[k:k + 20000]; co = zlib.compressobj(6, zlib.DEFLATED, -15); cd = co.compress(c) + co.flush()
            out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c))
        return out + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    rng = random.Random(7)
    ok = err = fmt_checked = 0
    for it in range(60):
        kind = it % 5
        if kind == 0:
            b = base[:rng.randint(0, len(base))]
        elif kind == 1:
            b = bytearray(base)
            for _ in range(rng.randint(1, 5)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        else:
            p = bytearray(payload)
            if kind == 2:
                for _ in range(rng.randint(1, 6)):
                    p[rng.randrange(len(p))] = rng.randrange(256)
            elif kind == 3:
                p = p[:rng.randint(0, len(p))]
            else:
                struct.pack_into("<i", p, rng.randrange(len(p) - 4), rng.choice([-1, 0, 1, 2 ** 31 - 1, -2 ** 31, 70000, 3]))
            b = bgzf(bytes(p))
        path = str(tmp_path / "m.bam"); open(path, "wb").write(bytes(b))
        try:
            nb.bam_dump_groups(path, str(tmp_path / "g.tsv"), force_bam_paired=bool(it & 1), num_cores=2)
            ok += 1
        except nb.NbError:
            err += 1
            continue
        # the rows stage's formatter walks the same damaged records: same values as the field parser, nothing read past a record
        os.environ["NB_BAM_DUMP_ROWFMT"] = "1"
        try:
            nb.bam_dump_groups(path, str(tmp_path / "r.tsv"), force_bam_paired=bool(it & 1), num_cores=2)
        finally:
            del os.environ["NB_BAM_DUMP_ROWFMT"]
        fl = open(tmp_path / "g.tsv", "rb").read().split(b"\n")[:-1]; rl = open(tmp_path / "r.tsv", "rb").read().split(b"\n")[:-1]
        if len(fl) == len(rl) and all(l.count(b"\t") == 40 for l in fl):      # (no tab / newline inside a damaged value)
            for a, b2 in zip(fl, rl):
                f = a.split(b"\t")[3:]
                assert b2 == b"\t".join(f[i] for i in range(38) if i not in (1, 15))
            fmt_checked += len(fl)
    assert ok + err == 60 and err >= 5 and fmt_checked > 1000


@pytest.mark.parametrize("force_paired", [False, True])
@pytest.mark.parametrize("window_kb", [64, 300])
def test_streaming_windows_equal_the_whole_file_readers(tmp_path, monkeypatch, force_paired, window_kb):
    """The producer reads the BAM window by window with bounded memory (a window holds back its last complete run and what
    follows it); whatever the window size, the groups must be those of the serial whole-file readers — including the
    end-of-file quirks (last buffer unsorted, last group never sent)."""
    L = synth.SynthLibrary(seed=7, n_fam=20, n_all=5)
    bam = make_bam(str(tmp_path / "w.bam"), L, n_groups=3000, seed=11)
    a, b = str(tmp_path / "win.tsv"), str(tmp_path / "ser.tsv")
    monkeypatch.setenv("NB_BAM_WINDOW_KB", str(window_kb))
    nb.bam_dump_groups(bam, a, force_bam_paired=force_paired, num_cores=4)
    monkeypatch.delenv("NB_BAM_WINDOW_KB")
    monkeypatch.setenv("NB_BAM_SERIAL_GROUPING", "1")
    nb.bam_dump_groups(bam, b, force_bam_paired=force_paired, num_cores=4)
    wa, se = open(a, "rb").read(), open(b, "rb").read()
    assert wa == se and wa.count(b"\n") > (300 if force_paired else 3000)
    assert os.path.getsize(bam) > 4 * window_kb * 1024 // 4      # several windows of compressed data


def test_feeder_rejects_size_fields_that_point_outside_the_file(tmp_path):
    """xlen / bsize / isize of a BGZF block and l_read_name / n_cigar / l_seq of a record are trusted nowhere;
    an unterminated Z aux value at the end of a record is not a string."""
    import struct
    import zlib
    L = synth.SynthLibrary(seed=1234, n_fam=20, n_all=5)
    base = open(make_bam(str(tmp_path / "t.bam"), L, n_groups=30), "rb").read()
    out = str(tmp_path / "g.tsv")

    def fails(b):
        path = str(tmp_path / "d.bam"); open(path, "wb").write(bytes(b))
        with pytest.raises(nb.NbError):
            nb.bam_dump_groups(path, out, num_cores=2)
    b = bytearray(base); struct.pack_into("<H", b, 10, 0xFFFF); fails(b)            # xlen runs past the file
    b = bytearray(base); struct.pack_into("<H", b, 16, 5); fails(b)                 # bsize smaller than its own header
    bs = struct.unpack_from("<H", base, 16)[0] + 1
    b = bytearray(base); struct.pack_into("<I", b, bs - 4, 1 << 20); fails(b)       # isize beyond 64 KiB
    # record fields inside intact BGZF blocks
    payload, i = b"", 0
    while i < len(base):
        xlen = struct.unpack_from("<H", base, i + 10)[0]; bsz = struct.unpack_from("<H", base, i + 16)[0] + 1
        payload += zlib.decompress(base[i + 12 + xlen:i + bsz - 8], -15); i += bsz

    def bgzf(data):
        o = b""
        for k in range(0, len(data), 20000):
            c = data[k:k + 20000]; co = zlib.compressobj(6, zlib.DEFLATED, -15); cd = co.compress(c) + co.flush()
            o += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(c) & 0xFFFFFFFF, len(c))
        return o + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    l_text = struct.unpack_from("<I", payload, 4)[0]; p = 8 + l_text; n_ref = struct.unpack_from("<I", payload, p)[0]; p += 4
    for _ in range(n_ref):
        p += 4 + struct.unpack_from("<I", payload, p)[0] + 4
    rec = p + 4                                                                     # first record's fixed fields
    q = bytearray(payload); struct.pack_into("<i", q, rec + 16, 1 << 28); fails(bgzf(bytes(q)))      # l_seq
    q = bytearray(payload); struct.pack_into("<H", q, rec + 12, 0xFFFF); fails(bgzf(bytes(q)))      # n_cigar
    # unterminated Z value as the last aux field of the first record: it is not a CB / UB string, nothing is read past the record
    blk = struct.unpack_from("<I", payload, p)[0]
    q = bytearray(payload)
    end = p + 4 + blk
    z = bytes(q[p + 4:end]).rfind(b"\0")                                            # the NUL of the last Z value
    q[p + 4 + z] = ord("X")
    path = str(tmp_path / "z.bam"); open(path, "wb").write(bgzf(bytes(q)))
    try:
        nb.bam_dump_groups(path, out, num_cores=2)                                  # groups (the field is ignored) or a loud error, never a crash
    except nb.NbError:
        pass


@pytest.mark.parametrize("force_paired", [False, True])
@pytest.mark.parametrize("window_kb", [64, 700, 1 << 20])
def test_parallel_record_walk_and_run_detection(tmp_path, monkeypatch, force_paired, window_kb):
    """The windowed producer walks a window's record chain in segments (guessed starts, accepted only when every segment
    begins where the previous one ended) and finds the kept records / UMI runs with thread-local counts.  Both are normally
    used on windows of megabytes only; NB_BAM_PAR_MIN_* forces them on this small file, at several window sizes, and the
    dump must stay byte-identical to the serial readers' (which define the semantics)."""
    L = synth.SynthLibrary(seed=7, n_fam=20, n_all=5)
    bam = make_bam(str(tmp_path / "w.bam"), L, n_groups=3000, seed=11)
    ser, par = str(tmp_path / "ser.tsv"), str(tmp_path / "par.tsv")
    monkeypatch.setenv("NB_BAM_SERIAL_GROUPING", "1")
    nb.bam_dump_groups(bam, ser, force_bam_paired=force_paired, num_cores=5)
    monkeypatch.delenv("NB_BAM_SERIAL_GROUPING")
    monkeypatch.setenv("NB_BAM_PAR_MIN_BYTES", "1")
    monkeypatch.setenv("NB_BAM_PAR_MIN_RECS", "1")
    monkeypatch.setenv("NB_BAM_WINDOW_KB", str(window_kb))
    nb.bam_dump_groups(bam, par, force_bam_paired=force_paired, num_cores=5)
    a, b = open(ser, "rb").read(), open(par, "rb").read()
    assert a == b and a.count(b"\n") > (300 if force_paired else 3000)


def test_row_formatter_equals_the_field_parser(tmp_path, monkeypatch):
    """The rows stage writes a record's 36 reported values with its own formatter (one pass over the aux block, fields
    grouped by the two bytes htslib's aux lookup reads); the dump's field parser and oracle/bam_ref.py define the values.
    Records here carry every aux type, tags that several fields share ("MA": MATE_REVERSE / MATE_UNMAPPED / MAPQ /
    MATE_POS; "RE": REVERSE / RE; "QN": QNAME), repeated tags (the first one counts), negative positions and odd flags."""
    import random
    from synth import bamio
    rng = random.Random(11)
    nt = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    extras = [("MA", "Z", "shared-by-four"), ("RE", "Z", "both"), ("RE", "A", "E"), ("QN", "Z", "renamed"), ("QN", "i", 5), ("SE", "Z", "seq-len-too"), ("PA", "Z", "p"),
              ("NH", "Z", "as-string"), ("NH", "i", 3), ("HI", "C", 200), ("AS", "s", -300), ("nM", "S", 65535), ("fx", "Z", "feat;ure"), ("TX", "Z", "tx,+,1M"), ("AN", "H", "1AE3"),
              ("xx", "B", ("s", [1, -2, 3])), ("GN", "B", ("C", [1, 2])), ("GN", "Z", "behind-an-array"), ("yy", "f", 1.5), ("zz", "I", 4000000000), ("cc", "c", -5), ("UY", "Z", ""), ("SK", "Z", "user-SK")]
    recs = []
    for g in range(400):
        umi, cb = nt(12), nt(16) + "-1"
        for k in range(rng.randint(1, 4)):
            tags = [("CB", "Z", cb), ("UB", "Z", umi)] + rng.sample(extras, rng.randint(0, 6))
            rng.shuffle(tags)
            flags = rng.choice([[0], [16], [4], [256], [512 | 16], [1024], [2048 | 128], [0xFFE], [1 | 64 | 32, 1 | 128 | 16], [1 | 2 | 128, 1 | 2 | 64], [1 | 8 | 64, 1 | 4 | 128], [0xFFF, 0xFFF]])
            for flag in flags:      # records flagged as paired come as two neighbours with one QNAME (anything else is filtered out)
                L = rng.choice([91, 124, 30, 1])
                recs.append(bamio.encode_record("q%d_%d" % (g, k), flag, nt(L), bytes(rng.randrange(2, 41) for _ in range(L)), tags, refid=rng.choice([-1, 0]), pos=rng.choice([-1, 0, 7, 2 ** 31 - 1]),
                                                mapq=rng.randrange(256), next_refid=rng.choice([-1, 0]), next_pos=rng.choice([-1, 7, 99, 123456789]), tlen=rng.choice([0, -250, 2 ** 31 - 1, -2 ** 31])))
    bam = bamio.write_bam(str(tmp_path / "f.bam"), recs, block_bytes=20000)
    keep = [i for i in range(38) if i not in (1, 15)]
    for force_paired in (False, True):
        a, b = str(tmp_path / "fields.tsv"), str(tmp_path / "rows.tsv")
        monkeypatch.delenv("NB_BAM_DUMP_ROWFMT", raising=False)
        nb.bam_dump_groups(bam, a, force_bam_paired=force_paired, num_cores=3)
        monkeypatch.setenv("NB_BAM_DUMP_ROWFMT", "1")
        nb.bam_dump_groups(bam, b, force_bam_paired=force_paired, num_cores=3)
        fields = [l.rstrip("\n").split("\t")[3:] for l in open(a, encoding="latin1")]
        rows = [l.rstrip("\n").split("\t") for l in open(b, encoding="latin1")]
        ref = [it["f"] for g in bam_ref.groups_of(bam_ref.read_bam(bam), force_paired) for it in g]
        assert len(rows) == len(fields) == len(ref) > 300
        for r, f, o in zip(rows, fields, ref):
            assert r == [f[i] for i in keep] == [o[i] for i in keep]
        assert any(r[0] == "renamed" for r in rows) and any(r[2] == r[7] == r[11] == r[13] == "shared-by-four" for r in rows) and any(r[1] == "both" for r in rows)
