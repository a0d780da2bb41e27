"""GPU end-to-end test of BAM mode: nb_process_bam (feeder -> scoped device batches -> gzip TSV)
against the Python restatement of process::bam::process on the CPU oracle, compared on the deterministic projection
of SURVEY.md Appendix F: header, per (UMI, CB) scope the multiset of (features, score) rows, the zero-row count, and for
every row the per-read_key filter columns of the pair it shows."""
import collections
import gzip
import json

import pytest

import nimble_aligner_b200 as nb
import oracle as orc
import synth
from oracle import bam_ref
from tests.bamcases import make_bam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chem,force_paired,trim", [("unstranded", False, None), ("fiveprime", False, "30:0.5"), ("threeprime", False, None), ("unstranded", True, None)])
def test_process_bam_matches_restatement(tmp_path, chem, force_paired, trim):
    L = synth.SynthLibrary(seed=1234, n_fam=40, n_all=5, group_on="family", num_mismatches=1)
    obj = L.to_json_obj()
    lib_json = tmp_path / "lib.json"
    lib_json.write_text(json.dumps(obj))
    bam = make_bam(str(tmp_path / "t.bam"), L, n_groups=400, paired_fraction=0.3 if force_paired else 0.15)
    out = tmp_path / "out.tsv.gz"
    nb.process_bam(bam, [str(lib_json)], [str(out)], chem, trim=trim, num_cores=4, force_bam_paired=force_paired)
    lines = gzip.open(out, "rt", encoding="latin1").read().split("\n")
    assert lines[-1] == ""
    lines = lines[:-1]
    # ---- reference side
    ocfg, oref = orc.parse_reference_library(obj, chem)
    if trim:
        ocfg["trim_target_length"], ocfg["trim_strictness"] = int(trim.split(":")[0]), float(trim.split(":")[1])
    groups = bam_ref.groups_of(bam_ref.read_bam(bam), force_paired)
    want = bam_ref.align_groups(orc.Oracle(ocfg, oref), groups)
    assert sum(1 for w in want if w["rows"]) > 20
    assert lines[0] == bam_ref.header_line()
    hdr = lines[0].split("\t")
    col = {n: i for i, n in enumerate(hdr)}
    assert len(hdr) == 2 + 36 + 36 + 10
    got = collections.OrderedDict()
    for l in lines[1:]:
        f = l.split("\t")
        assert len(f) == len(hdr)
        umi = f[col["r2_UB"]] or f[col["r2_UR"]]
        got.setdefault((umi, f[col["r2_CB"]][:-2]), []).append(f)
    exp = collections.OrderedDict()
    for g, w in zip(groups, want):
        if w["rows"]:
            assert w["key"] not in exp
            exp[w["key"]] = (g, w)
    assert list(got.keys()) == list(exp.keys())          # scopes with >= 1 callset, in file order; others emit nothing
    for key, rows in got.items():
        g, w = exp[key]
        nonzero = sorted((r[0], int(r[1])) for r in rows if r[0] != "")
        assert nonzero == sorted((",".join(cs), n) for cs, n in w["rows"]), key
        assert sum(1 for r in rows if r[0] == "") == w["n_zero_rows"], key
        by_q = {p["qname"]: p for p in w["pairs"]}
        md = {it["f"][0]: None for it in g}
        seen = set()
        for r in rows:
            q = r[col["r2_QNAME"]]
            assert q == r[col["r1_QNAME"]] and q in by_q and q not in seen
            seen.add(q)
            p = by_q[q]
            assert r[col["r2_filter_forward"]] == orc.REASONS[p["fr1"]] and r[col["r1_filter_forward"]] == orc.REASONS[p["fr2"]], (key, q)
            assert int(r[col["r2_forward_score"]]) == p["score1"] and int(r[col["r1_forward_score"]]) == p["score2"]
            assert r[col["triage_reason"]] == orc.REASONS[p["triage"]]
            assert r[col["r1_filter_reverse"]] == "None" and r[col["aligndirection"]] == "None"
            if r[0] != "":   # the representative shown for a callset row really resolved to that callset
                assert p["callset"] is not None and ",".join(p["callset"]) == r[0]
            # metadata columns are the record's 36 reported fields: sequence slot under r2_, mate slot under r1_
            j = [i for i, it in enumerate(g) if it["f"][0] == q]
            seq_slot, mate_slot = g[j[0]], g[j[1]]
            assert "\t".join(r[2:38]) == bam_ref.data_values(mate_slot["f"]) and "\t".join(r[38:74]) == bam_ref.data_values(seq_slot["f"])
    # the nimble CLI drives the same path (src/bin/cli.yml arguments)
    import subprocess, os
    cli = os.path.join(os.path.dirname(nb.SO_PATH), "nimble")
    out2 = tmp_path / "out2.tsv.gz"
    args = [cli, "-r", str(lib_json), "-o", str(out2), "-i", bam, "-c", "4", "--strand_filter", chem] + (["-t", trim] if trim else []) + (["-p"] if force_paired else [])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    assert gzip.open(out2, "rb").read() == gzip.open(out, "rb").read()
