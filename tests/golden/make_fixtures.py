"""Copies the reference's small test DATA fixtures (library JSONs + FASTQs; no source code) into tests/golden/ref/
and writes expected.json holding the literal expectations of the reference's own tests, so that the oracle and the
GPU path can be pinned on boxes where /root/reference does not exist.  Run here (container) only:
    python tests/golden/make_fixtures.py
Sources: /root/reference/tests/test-sequences/{libraries,reads}/, expectations transcribed from
tests/basic-cases.rs:59-71,95-107,131-143,165-177,201-213,237-249,273-277,300-304 and tests/mismatch.rs:30,57."""
import json, os, shutil
REF = "/root/reference/tests/test-sequences"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref")
os.makedirs(os.path.join(OUT, "libraries"), exist_ok=True)
os.makedirs(os.path.join(OUT, "reads"), exist_ok=True)
for f in sorted(os.listdir(os.path.join(REF, "libraries"))):
    shutil.copy(os.path.join(REF, "libraries", f), os.path.join(OUT, "libraries", f))
for f in sorted(os.listdir(os.path.join(REF, "reads"))):
    if f.endswith(".fastq"):
        shutil.copy(os.path.join(REF, "reads", f), os.path.join(OUT, "reads", f))
A4 = ["A02-0", "A02-1", "A02-2", "A02-LC"]
basic01 = [[A4, 1], [["A02-0", "A02-LC"], 1], [["A02-1"], 1]]
basic2 = [[A4, 1], [["A02-0", "A02-LC"], 1], [["A02-1"], 2]]
grp = [[["g1"], 1], [["g1", "g2"], 1], [["g2"], 1]]
expected = {
    "get_calls": [  # library, reads, num_mismatches, group_on_test_column, expected sorted [(callset, count)]
        {"src": "tests/basic-cases.rs:44-74", "lib": "basic.json", "reads": "basic.fastq", "mm": 0, "group": False, "expect": basic01},
        {"src": "tests/basic-cases.rs:78-110", "lib": "basic.json", "reads": "basic.fastq", "mm": 1, "group": False, "expect": basic01},
        {"src": "tests/basic-cases.rs:114-146", "lib": "basic.json", "reads": "basic.fastq", "mm": 2, "group": False, "expect": basic2},
        {"src": "tests/basic-cases.rs:150-180", "lib": "basic-rev.json", "reads": "basic.fastq", "mm": 0, "group": False, "expect": basic01},
        {"src": "tests/basic-cases.rs:184-216", "lib": "basic-rev.json", "reads": "basic.fastq", "mm": 1, "group": False, "expect": basic01},
        {"src": "tests/basic-cases.rs:220-252", "lib": "basic-rev.json", "reads": "basic.fastq", "mm": 2, "group": False, "expect": basic2},
        {"src": "tests/basic-cases.rs:256-280", "lib": "basic.json", "reads": "basic.fastq", "mm": 0, "group": True, "expect": grp},
        {"src": "tests/basic-cases.rs:285-307", "lib": "basic.json", "reads": "basic.fastq", "mm": 0, "group": True, "expect": grp},
        {"src": "tests/mismatch.rs:11-33", "lib": "mismatch.json", "reads": "mismatch.fastq", "mm": 0, "group": False, "expect": [[["gene"], 1]]},
        {"src": "tests/mismatch.rs:37-60", "lib": "mismatch.json", "reads": "mismatch.fastq", "mm": 1, "group": False, "expect": [[["gene"], 2]]},
    ],
    "group_column": ["g1", "g1", "g2", "g2", "g2", "g2", "g1", "g1", "g1", "g1"],  # tests/basic-cases.rs:29-36
}
json.dump(expected, open(os.path.join(HERE, "expected.json"), "w"), indent=1)
print("wrote", OUT)
