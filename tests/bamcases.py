"""Synthetic 10x-style BAMs for the BAM-mode tests (shared by the CPU feeder test and the GPU end-to-end test)."""
import numpy as np

import synth
from synth import bamio


def make_bam(path, L, n_groups=300, seed=2345, read_len=91, paired_fraction=0.15, tso_fraction=0.1):
    """Writes a BAM whose records are contiguous by UMI (as sorted_bam_reader.rs requires) and exercises: unpaired 10x
    records (dummy mates), true pairs (first/last in template, either order), REVERSE records, 124-base records (TSO
    clip), UB missing -> UR, records without CB (skipped), the AAAAAAAAAA UMI (skipped), several CBs inside one UMI
    (the CB sort), Q2 tails."""
    rng = np.random.default_rng(seed)
    u = synth.umi_reads(L, 0, n_groups, seed=seed, L=read_len)
    bases = u["bases"][: u["n_reads"] * read_len].reshape(-1, read_len)
    qual = u["qual"][: u["n_reads"] * read_len].reshape(-1, read_len)
    seqs = L.sequences()
    recs = []
    nt = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    start = np.concatenate([[0], np.cumsum(u["sizes"])])
    qn = 0
    for g in range(n_groups):
        umi = nt(12)
        cbs = [nt(16) + "-1" for _ in range(1 + int(rng.integers(0, 3) == 0))]    # sometimes two cells share the UMI string
        order = list(range(int(start[g]), int(start[g + 1])))
        for ri in order:
            cb = cbs[int(rng.integers(len(cbs)))]
            s = bytes(bases[ri]).decode(); q = bytes(qual[ri])
            flag = 16 if rng.random() < 0.3 else 0
            if rng.random() < tso_fraction:   # 124-base record: 13 non-biological bases at the 5' end of the original read
                pad = nt(124 - read_len); padq = bytes([30] * (124 - read_len))
                s, q = (s + pad[:124 - read_len], q + padq) if flag & 16 else (pad + s, padq + q)
                s = s[:124] if len(s) >= 124 else s + nt(124 - len(s)); q = (q + bytes([30] * 124))[:124]
            tags = [("CB", "Z", cb), ("CR", "Z", cb[:-2]), ("CY", "Z", "F" * 16), ("UR", "Z", umi), ("UY", "Z", "F" * 12), ("NH", "i", 1), ("RE", "A", "E"), ("GN", "Z", "GENE%d" % (g % 7))]
            if rng.random() < 0.9:
                tags.append(("UB", "Z", umi))
            if rng.random() < 0.03:
                tags = [t for t in tags if t[0] != "CB"]                          # no CB -> skipped
            name = "q%07d" % qn; qn += 1
            if rng.random() < paired_fraction:   # a true pair: mate = reverse complement of a downstream window of the same transcript
                t = seqs[int(rng.integers(len(seqs)))]
                st = int(rng.integers(0, len(t) - 300))
                a = t[st:st + read_len]; bseq = t[st + 150:st + 150 + read_len]
                first_flag, last_flag = 1 | 2 | 64 | 32, 1 | 2 | 128 | 16
                ra = bamio.encode_record(name, first_flag, a, bytes([36] * read_len), tags, refid=0, pos=1000 + st, next_refid=0, next_pos=1150 + st, tlen=150 + read_len)
                rb = bamio.encode_record(name, last_flag, bseq, bytes([36] * read_len), tags, refid=0, pos=1150 + st, next_refid=0, next_pos=1000 + st, tlen=-(150 + read_len))
                recs += [rb, ra] if rng.random() < 0.5 else [ra, rb]
            else:
                recs.append(bamio.encode_record(name, flag, s, q, tags, refid=0 if flag else -1, pos=int(rng.integers(1, 10 ** 6))))
        if g % 50 == 7:   # a block of the whitelisted UMI in between: skipped entirely
            for _ in range(3):
                recs.append(bamio.encode_record("w%07d" % qn, 0, nt(read_len), bytes([36] * read_len), [("CB", "Z", cbs[0]), ("UB", "Z", "AAAAAAAAAA"), ("UR", "Z", "AAAAAAAAAA")])); qn += 1
    bamio.write_bam(path, recs, block_bytes=20000)
    return path
