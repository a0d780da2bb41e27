"""CPU-only tests of the product's host side: the C-ABI library loads and exports every declared symbol, the library
loader and the index builder agree with the oracle (independent implementations), and compute entry points fail
loudly without a GPU.  No device compute here."""
import ctypes
import os
import random
import re

import pytest

import nimble_aligner_b200 as nb
import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "ref")
LIBS = ["basic.json", "basic-rev.json", "mismatch.json", "strandedness.json", "reference-library-correct.json",
        "reference-library-rna.json", "reference-library-mixed-case-rna.json", "reference-library-no-rna-bases.json"]


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nimble_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) > 30
    so = ctypes.CDLL(nb.SO_PATH)
    missing = [s for s in declared if not hasattr(so, s)]
    assert not missing, missing
    assert nb.lib().nb_version().decode().startswith("nimble_b200")
    assert nb.lib().nb_reason_str(10).decode() == "Low Entropy"   # src/align.rs:68


@pytest.mark.parametrize("name", LIBS)
def test_library_loader_matches_oracle_loader(name):
    path = os.path.join(G, "libraries", name)
    cfg, ref = orc.get_reference_library(path, "fiveprime")
    c, lib = nb.get_reference_library(path, "fiveprime")
    assert lib.headers == ref.headers
    for i in range(len(ref.headers)):
        assert lib.column(i) == ref.columns[i]
    assert (lib.group_on, lib.sequence_name_idx, lib.sequence_idx) == (ref.group_on, ref.sequence_name_idx, ref.sequence_idx)
    for k in ("score_percent", "score_threshold", "num_mismatches", "discard_multi_hits", "max_hits_to_report", "intersect_level",
              "trim_target_length", "trim_strictness", "score_filter", "reference_genome_size"):
        assert getattr(c, k) == cfg[k], k
    assert bool(c.discard_multiple_matches) == cfg["discard_multiple_matches"] and bool(c.require_valid_pair) == cfg["require_valid_pair"]
    assert c.strand_filter == nb.CHEM["fiveprime"] and c.discard_nonzero_mismatch == 0
    seqs, names = nb.get_reference_sequence_data(lib)   # src/utils.rs:121-190
    assert names == ref.columns[ref.sequence_name_idx] and seqs == ref.columns[ref.sequence_idx]


@pytest.mark.parametrize("ensure_ascii", [True, False])
def test_library_loader_strings_with_escapes_and_long_runs(ensure_ascii):
    """The JSON reader appends unescaped runs wholesale and moves strings out of the parsed tree: escapes in the middle of
    a run, \\u escapes incl. surrogate pairs, raw UTF-8, empty strings and a 100 kb sequence must come out as serde_json
    (here: Python's json) reads them, and every row must get its revcomp twin (src/reference_library.rs:130-153)."""
    import json
    rng = random.Random(5)
    names = ['plain', 'quote"inside', 'back\\slash', 'tab\there', '\u00e9-accent', 'emoji-\U0001F9EC-dna', 'slash/', 'ctl\x01', '', 'trailing\\']
    seqs = ["".join(rng.choice("ACGTUacgtuNn") for _ in range(n)) for n in (31, 40, 100_000, 64, 1, 33, 257, 1024, 30, 4096)]
    groups = ['g"1', 'g"1', '', 'g\\2', 'g\\2', 'g3', 'g3', '\u00e9', '\u00e9', '']
    cfg = dict(score_percent=0.5, score_filter=25, score_threshold=50, num_mismatches=1, discard_multiple_matches=False, require_valid_pair=True,
               discard_multi_hits=0, intersect_level=1, max_hits_to_report=4, group_on="grp", trim_target_length=40, trim_strictness=0.9)
    obj = [cfg, {"headers": ["sequence_name", "grp", "sequence"], "columns": [names, groups, seqs]}]
    text = json.dumps(obj, ensure_ascii=ensure_ascii, indent=1 if ensure_ascii else None)
    lib = nb.Library.from_text(text, "unstranded")
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a", "N": "N", "n": "N"}
    want_names, want_groups, want_seqs = [], [], []
    for n, g, sq in zip(names, groups, seqs):
        sq = sq.replace("U", "T").replace("u", "t")
        want_names += [n, n + "\u00a7rev"]; want_groups += [g, g]
        want_seqs += [sq, "".join(comp[c] for c in reversed(sq))]
    assert lib.column(0) == want_names and lib.column(1) == want_groups and lib.column(2) == want_seqs
    assert lib.headers == ["sequence_name", "grp", "sequence"] and lib.group_on == 1
    c = lib.config
    assert (c.score_percent, c.num_mismatches, c.max_hits_to_report, bool(c.require_valid_pair), c.reference_genome_size) == (0.5, 1, 4, True, 10)


def test_library_loader_survives_mutated_input():
    """A library file is outside input: truncations, flipped bytes, stray structural characters and deletions must end in a
    parsed library or an NbError (the reference panics with a message), never in a crash; bytes that are not UTF-8 are
    refused like fs::read_to_string does (src/reference_library.rs:21)."""
    base = open(os.path.join(G, "libraries", "basic.json"), "rb").read()
    rng = random.Random(1)
    parsed = rejected = 0
    for _ in range(600):
        b = bytearray(base)
        k = rng.randint(0, 3)
        if k == 0:
            b = b[:rng.randint(0, len(b))]
        elif k == 1:
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        elif k == 2:
            i = rng.randrange(len(b)); b[i:i] = bytes(rng.choice(b'{}[]",:\\u0123 \n') for _ in range(rng.randint(1, 6)))
        else:
            i = rng.randrange(len(b)); del b[i:min(len(b), i + rng.randint(1, 50))]
        try:
            lib = nb.Library.from_text(bytes(b), "unstranded")
        except nb.NbError:
            rejected += 1
            continue
        parsed += 1
        bytes(b).decode()                                    # whatever was accepted is UTF-8
        assert lib.n_rows % 2 == 0 and len(lib.group_names()) <= lib.n_rows
    assert parsed > 20 and rejected > 200


@pytest.mark.parametrize("name", ["reference-library-missing-fields.json", "reference-library-types-broken.json", "reference-library-broken-format.json"])
def test_library_loader_rejects_what_the_reference_panics_on(name):   # src/reference_library.rs:256-300
    with pytest.raises(nb.NbError):
        nb.get_reference_library(os.path.join(G, "libraries", name))
    with pytest.raises(nb.NbError):
        nb.get_reference_library(os.path.join(G, "libraries", "does-not-exist.json"))


def test_sanity_check_align_config():   # src/reference_library.rs:209-226
    _, lib = nb.get_reference_library(os.path.join(G, "libraries", "basic.json"))
    for bad in (dict(score_percent=1.5), dict(score_percent=-0.1), dict(trim_strictness=2.0), dict(score_filter=-1)):
        with pytest.raises(nb.NbError):
            lib.set_config(lib.config.copy(**bad))
    lib.set_config(lib.config.copy(score_percent=1.0, trim_strictness=0.0))


def _index_parity(names, seqs, threads=3):
    cfg = dict(score_percent=0.1, score_threshold=50, num_mismatches=0, discard_multiple_matches=False, require_valid_pair=False,
               discard_multi_hits=0, max_hits_to_report=5, intersect_level=0, strand_filter="none", trim_target_length=15, trim_strictness=0.5)
    o = orc.Oracle(cfg, orc.Reference(0, ["sequence_name", "sequence"], [names, seqs], 0, 1))
    ix = nb.Index.from_sequences(seqs, threads)
    so, sp = o.index_stats(), ix.stats()
    for k in so:
        assert so[k] == sp[k], k
    assert o.index_dump() == ix.dump()


@pytest.mark.parametrize("name", ["basic.json", "basic-rev.json", "mismatch.json", "strandedness.json"])
def test_index_matches_oracle_on_reference_fixtures(name):
    _, ref = orc.get_reference_library(os.path.join(G, "libraries", name), "none")
    _index_parity(ref.columns[ref.sequence_name_idx], ref.columns[ref.sequence_idx])


def test_index_matches_oracle_on_synthetic_family_library():
    import synth
    L = synth.SynthLibrary(seed=7, n_fam=12, n_all=5)
    cfg, ref = orc.parse_reference_library(L.to_json_obj(), "none")
    _index_parity(ref.columns[ref.sequence_name_idx], ref.columns[ref.sequence_idx], threads=4)


def test_index_edge_cases_cycles_repeats_short_and_non_acgt():
    rnd = random.Random(5)
    rs = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    core = rs(60)
    seqs = ["A" * 50,                    # homopolymer: a k-mer that is its own successor (pure 1-cycle)
            "AC" * 40, "CA" * 40,        # 2-cycle entered at different phases
            "ACG" * 30,                  # 3-cycle
            rs(29), "", rs(30),          # shorter than k contributes nothing; exactly k = one k-mer
            core + rs(40), rs(35) + core + rs(20), core,   # shared segments -> colour changes and forks
            "ACGTNNNNACGTRYKM" * 5,      # non-ACGT -> A
            (rs(45) * 3)[:120],          # tandem repeat longer than k: cycle with a tail
            "acgtacgtacgtacgtacgtacgtacgtacgtacgtacgt"]  # lowercase
    names = ["s%d" % i for i in range(len(seqs))]
    _index_parity(names, seqs, threads=2)
    _index_parity(names, seqs, threads=1)


def test_group_names_follow_natural_lexical_order():
    cfg = nb.Config(score_percent=0.1, max_hits_to_report=5)
    names = ["b10", "B9", "a1", "A02-LC", "A02-2"]
    lib = nb.Library.from_columns(["sequence_name", "sequence"], [names, ["ACGT" * 10] * 5], 0, 0, 1, cfg)
    import functools
    assert lib.group_names() == sorted(names, key=functools.cmp_to_key(orc.natural_lexical_cmp))


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if nb.lib().nb_device_count() > 0:
        pytest.skip("GPU present")
    _, lib = nb.get_reference_library(os.path.join(G, "libraries", "basic.json"))
    ix = nb.build_index(lib, 2)
    with pytest.raises(nb.NbError) as e:
        nb.Context(ix, lib)
    assert e.value.code == -6 and "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nimble_aligner_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f


def test_index_save_load_roundtrip(tmp_path):
    _, lib = nb.get_reference_library(os.path.join(G, "libraries", "basic.json"))
    ix = nb.build_index(lib, 2)
    p = tmp_path / "basic.nbidx"
    ix.save(p)
    ix2 = nb.Index.load(p)
    assert ix2.stats() == ix.stats() and ix2.dump() == ix.dump()
    (tmp_path / "bad.nbidx").write_bytes(p.read_bytes()[:100])
    with pytest.raises(nb.NbError):
        nb.Index.load(tmp_path / "bad.nbidx")
    with pytest.raises(nb.NbError):
        nb.Index.load(tmp_path / "missing.nbidx")
    # the file ends in a checksum: a single flipped byte anywhere, an element count the file cannot hold, or a missing
    # tail is refused instead of being uploaded (a damaged node or edge id would send the kernels out of bounds)
    good = p.read_bytes()
    rng = random.Random(11)
    for _ in range(40):
        b = bytearray(good); i = rng.randrange(len(b)); b[i] ^= 1 << rng.randrange(8)
        (tmp_path / "bad.nbidx").write_bytes(bytes(b))
        with pytest.raises(nb.NbError):
            nb.Index.load(tmp_path / "bad.nbidx")
    for cut in (len(good) - 1, len(good) - 8, len(good) // 2):
        (tmp_path / "bad.nbidx").write_bytes(good[:cut])
        with pytest.raises(nb.NbError):
            nb.Index.load(tmp_path / "bad.nbidx")
    b = bytearray(good); b[40:48] = (2 ** 62).to_bytes(8, "little")   # first array's element count
    (tmp_path / "bad.nbidx").write_bytes(bytes(b))
    with pytest.raises(nb.NbError):
        nb.Index.load(tmp_path / "bad.nbidx")


def test_index_compare_and_gpu_builder_fails_loudly_without_a_gpu():
    rnd = random.Random(9)
    rs = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    core = rs(50)
    seqs = [core + rs(60), rs(20) + core, rs(90)]
    a, b = nb.Index.from_sequences(seqs, 1), nb.Index.from_sequences(seqs, 4)
    assert a.compare(b) == 0
    assert a.compare(nb.Index.from_sequences(seqs[:2] + [rs(90)], 1)) != 0
    if nb.lib().nb_device_count() == 0:
        with pytest.raises(nb.NbError) as e:
            nb.Index.from_sequences(seqs, 1, device=0)
        assert e.value.code == -6


def test_index_cache_keyed_by_the_library(tmp_path):
    """the drivers take the index from $NB_INDEX_CACHE/<key>.nbix when it is there.  The key follows the
    sequence column as the index sees it (case and non-ACGT folded) and nothing else; a hit needs no GPU; a damaged file is
    not a hit."""
    import json
    base = json.load(open(os.path.join(G, "libraries", "basic.json")))

    def lib_of(obj, name):
        p = tmp_path / name
        p.write_text(json.dumps(obj))
        return nb.get_reference_library(str(p))[1]
    lib = lib_of(base, "a.json")
    key = nb.Index.cache_key(lib)
    assert re.fullmatch(r"[0-9a-f]{32}", key) and key == nb.Index.cache_key(lib_of(base, "b.json"))
    # same index: case folds like DnaString::from_acgt_bytes (the key sees the sequence column after the library's own
    # reverse-complement rows were added, so a letter that complements differently is another index)
    txt = json.dumps(base)
    seqs = sorted(set(re.findall(r'"([ACGT]{40,})"', txt)), key=len, reverse=True)
    assert seqs
    folded = txt.replace(seqs[0], seqs[0].lower(), 1)
    lf = lib_of(json.loads(folded), "c.json")
    assert (nb.Index.cache_key(lf) == key) == (nb.build_index(lf, 1).compare(nb.build_index(lib, 1)) == 0)
    # another index: one base changed
    s0 = seqs[0]; changed = txt.replace(s0, s0[:10] + ("C" if s0[10] != "C" else "G") + s0[11:], 1)
    assert nb.Index.cache_key(lib_of(json.loads(changed), "d.json")) != key
    # a hit: the file written under the key is taken, no device involved
    cache = tmp_path / "cache"; cache.mkdir()
    host = nb.build_index(lib, 2)
    host.save(cache / (key + ".nbix"))
    hit = nb.Index.build_cached(lib, cache, device=0, threads=2)
    assert hit.compare(host) == 0 and hit.dump() == host.dump()
    os.environ["NB_INDEX_CACHE"] = str(cache)
    try:
        assert nb.Index.build_cached(lib, None, device=0, threads=2).compare(host) == 0       # the drivers' way: through the environment
    finally:
        del os.environ["NB_INDEX_CACHE"]
    if nb.lib().nb_device_count() == 0:
        # a miss (another library, or a damaged file) means building, which needs the device — loudly
        for obj, name in ((json.loads(changed), "e.json"),):
            with pytest.raises(nb.NbError) as e:
                nb.Index.build_cached(lib_of(obj, name), cache, device=0)
            assert e.value.code == -6
        good = (cache / (key + ".nbix")).read_bytes()
        (cache / (key + ".nbix")).write_bytes(good[:len(good) // 2])
        with pytest.raises(nb.NbError) as e:
            nb.Index.build_cached(lib, cache, device=0)
        assert e.value.code == -6
        assert sorted(x.name for x in cache.iterdir()) == [key + ".nbix"]                        # no temporary files left behind
