"""CPU test of the FASTQ feeder (/root/reference/src/parse/fastq.rs:21-43): the parallel plain-text
parser (mapped file cut into byte chunks, record starts guessed locally and verified against the sequential parse) and
the gzip path must hand over exactly the records a plain sequential FASTQ reader sees, whatever the chunk size, the thread
count, the line endings, or how adversarial the quality strings are."""
import gzip
import random

import pytest

import nimble_aligner_b200 as nb


def seq_parse(text):
    """The feeder's grammar, sequentially (bio::io::fastq accepts multi-line records): '@' header, sequence lines up to the
    '+' line, quality lines until as many quality characters as bases were read; blank lines between records are skipped."""
    lines = [l[:-1] if l.endswith("\r") else l for l in text.split("\n")]
    if lines and lines[-1] == "":
        lines.pop()
    out, i = [], 0
    while True:
        while i < len(lines) and lines[i] == "":
            i += 1
        if i >= len(lines):
            return out
        assert lines[i].startswith("@")
        i += 1
        seq = ""
        while not (lines[i] != "" and lines[i][0] == "+"):
            seq += lines[i]
            i += 1
        i += 1
        q = 0
        while q < len(seq):
            q += len(lines[i])
            i += 1
        assert q == len(seq)
        out.append(seq)


def make_fastq(rng, n, multiline=False, crlf=False, blanks=False, nasty_quals=True, final_newline=True):
    """Records of ragged length whose quality strings often START with '@' or '+' (the characters a careless
    re-synchronisation would trip over)."""
    nl = "\r\n" if crlf else "\n"
    parts = []
    for i in range(n):
        L = rng.choice([1, 30, 91, 150, 151, 250]) if rng.random() < 0.3 else 150
        s = "".join(rng.choice("ACGTN") for _ in range(L))
        q = "".join(rng.choice("@+IF#5") if nasty_quals else "I" for _ in range(L))
        if rng.random() < 0.3:
            q = rng.choice("@+") + q[1:]
        if multiline and L > 60 and rng.random() < 0.5:
            c = rng.randrange(1, L)
            rec = "@r%d%s%s%s%s%s+%s%s%s%s%s" % (i, nl, s[:c], nl, s[c:], nl, nl, q[:c], nl, q[c:], nl)
        else:
            rec = "@r%d some comment%s%s%s+%s%s%s" % (i, nl, s, nl, ("r%d" % i) if rng.random() < 0.2 else "", nl + q, nl)
        parts.append(rec)
        if blanks and rng.random() < 0.05:
            parts.append(nl)
    text = "".join(parts)
    if not final_newline:
        text = text.rstrip("\r\n")
    return text


@pytest.mark.parametrize("kw", [dict(), dict(crlf=True), dict(multiline=True), dict(blanks=True, final_newline=False), dict(multiline=True, crlf=True, blanks=True)])
@pytest.mark.parametrize("chunk,cores", [(4096, 1), (4096, 7), (20000, 4), (1 << 20, 3)])
def test_parallel_parser_equals_sequential(tmp_path, kw, chunk, cores):
    rng = random.Random(hash((chunk, cores, tuple(sorted(kw)))) & 0xFFFF)
    text = make_fastq(rng, 1500, **kw)
    p = tmp_path / "a.fastq"
    p.write_bytes(text.encode())
    out = tmp_path / "a.txt"
    nb.fastq_dump([p], out, num_cores=cores, chunk_bytes=chunk)
    assert out.read_text().split("\n")[:-1] == seq_parse(text)


def test_paired_streams_are_walked_in_lockstep(tmp_path):
    """R1 plain (parallel chunks) and R2 gzip (serial blocks) are cut at different record numbers; pairs must still line up."""
    rng = random.Random(7)
    t1, t2 = make_fastq(rng, 4000), make_fastq(rng, 4000, multiline=True)
    (tmp_path / "r1.fastq").write_bytes(t1.encode())
    with gzip.open(tmp_path / "r2.fastq.gz", "wb") as f:
        f.write(t2.encode())
    out = tmp_path / "p.txt"
    nb.fastq_dump([tmp_path / "r1.fastq", tmp_path / "r2.fastq.gz"], out, num_cores=6, chunk_bytes=8192)
    want = ["%s\t%s" % (a, b) for a, b in zip(seq_parse(t1), seq_parse(t2))]
    assert out.read_text().split("\n")[:-1] == want


def test_feeder_errors(tmp_path):
    """process::fastq::process panics on malformed / unequal inputs (src/process/fastq.rs:20-24): a negative status here."""
    rng = random.Random(3)
    good = make_fastq(rng, 300, nasty_quals=False)
    (tmp_path / "a.fastq").write_bytes(good.encode())
    (tmp_path / "short.fastq").write_bytes(make_fastq(random.Random(3), 299, nasty_quals=False).encode())
    (tmp_path / "trunc.fastq").write_bytes(good[:-40].encode() + b"garbage\n")
    (tmp_path / "noat.fastq").write_bytes(good.replace("@r150 ", "r150 ").encode())
    (tmp_path / "empty.fastq").write_bytes(b"")
    out = tmp_path / "o.txt"
    for r2 in ("short.fastq", "trunc.fastq", "noat.fastq"):
        for chunk in (4096, 1 << 20):
            with pytest.raises(nb.NbError) as e:
                nb.fastq_dump([tmp_path / "a.fastq", tmp_path / r2], out, num_cores=4, chunk_bytes=chunk)
            assert e.value.code == -3
    with pytest.raises(nb.NbError) as e:
        nb.fastq_dump([tmp_path / "noat.fastq", tmp_path / "a.fastq"], out, num_cores=4, chunk_bytes=4096)
    assert e.value.code == -3 and "R1" in str(e.value)
    with pytest.raises(nb.NbError):
        nb.fastq_dump([tmp_path / "missing.fastq"], out)
    nb.fastq_dump([tmp_path / "empty.fastq"], out)
    assert out.read_text() == ""
    nb.fastq_dump([tmp_path / "empty.fastq", tmp_path / "empty.fastq"], out, num_cores=2)
    assert out.read_text() == ""


@pytest.mark.parametrize("gz_threads", [1, 3])
@pytest.mark.parametrize("chunk_kb", [1, 7, 4096])
@pytest.mark.parametrize("kw", [dict(), dict(crlf=True, multiline=True, blanks=True), dict(final_newline=False)])
def test_gzip_streams(tmp_path, monkeypatch, chunk_kb, kw, gz_threads):
    """.gz input.  One thread for the file: it inflates the mapped file (csrc/inflate.hpp) into text chunks, a second one
    parses them.  More threads: csrc/pgunzip.hpp inflates on all of them and every worker parses its own chunk, the consumer
    the few lines at the junctions.  Either way lines and records straddle chunk borders (1 KiB / 4 KiB chunks: every few
    lines), the file may consist of several gzip members (bgzip, cat) cut anywhere, and the records must be those of the
    sequential parse of the plain text."""
    monkeypatch.setenv("NB_GZ_CHUNK_KB", str(chunk_kb))
    monkeypatch.setenv("NB_GZ_THREADS", str(gz_threads))
    rng = random.Random(chunk_kb * 31 + len(kw))
    t1, t2 = make_fastq(rng, 3000, **kw), make_fastq(rng, 3000, **kw)
    with gzip.open(tmp_path / "r1.fastq.gz", "wb", compresslevel=1) as f:
        f.write(t1.encode())
    b2 = t2.encode(); cuts = sorted(rng.randrange(len(b2)) for _ in range(5))
    (tmp_path / "r2.fastq.gz").write_bytes(b"".join(gzip.compress(b2[a:b], 9) for a, b in zip([0] + cuts, cuts + [len(b2)])))      # members cut anywhere, also inside lines
    out = tmp_path / "p.txt"
    nb.fastq_dump([tmp_path / "r1.fastq.gz", tmp_path / "r2.fastq.gz"], out, num_cores=4)
    assert out.read_text().split("\n")[:-1] == ["%s\t%s" % (a, b) for a, b in zip(seq_parse(t1), seq_parse(t2))]


def test_damaged_gzip_input_is_an_error_not_an_early_end(tmp_path):
    """A truncated or corrupted .fastq.gz must fail the job (gzread's short count looked like the end of the file)."""
    rng = random.Random(5)
    text = make_fastq(rng, 4000, nasty_quals=False).encode()
    good = gzip.compress(text, 6)
    out = tmp_path / "o.txt"
    cases = {"cut.fastq.gz": good[:len(good) // 2], "trailer.fastq.gz": good[:-3], "flip.fastq.gz": good[:5000] + bytes([good[5000] ^ 0x55]) + good[5001:],
             "crc.fastq.gz": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:]}
    for name, b in cases.items():
        (tmp_path / name).write_bytes(b)
        for cores in (1, 2, 5):                                      # the serial reader and the parallel one
            with pytest.raises(nb.NbError) as e:
                nb.fastq_dump([tmp_path / name], out, num_cores=cores)
            assert e.value.code == -3, name
    assert any("gzip" in str(x) for x in [e.value])
    (tmp_path / "ok.fastq.gz").write_bytes(good + b"\0" * 100)        # padding behind the last member is ignored (gzip -d does the same)
    nb.fastq_dump([tmp_path / "ok.fastq.gz"], out, num_cores=2)
    assert out.read_text().split("\n")[:-1] == seq_parse(text.decode())
