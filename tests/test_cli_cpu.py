"""CPU test of the `nimble` CLI front end (csrc/nimble_main.cpp): the argument handling of /root/reference/src/bin/cli.yml
and src/bin/main.rs:12-162 — required arguments, --cores / --strand_filter / --trim validation, the dispatch on the input's
extension — plus the --devices list this implementation adds.  Every failure is the reference's: a message and exit code
101 (a Rust panic).  None of these cases reaches the GPU."""
import os
import subprocess

import pytest

import nimble_aligner_b200 as nb

CLI = os.path.join(os.path.dirname(nb.SO_PATH), "nimble")


def run(*args):
    p = subprocess.run([CLI, *args], capture_output=True, text=True)
    return p.returncode, p.stdout + p.stderr


@pytest.fixture(scope="module", autouse=True)
def built():
    nb.lib()                      # builds the library and the CLI when they are stale
    assert os.path.exists(CLI)


def test_help_and_version():
    rc, out = run("--help")
    assert rc == 0 and "USAGE: nimble" in out
    rc, out = run("-V")
    assert rc == 0 and out.startswith("nimble 0.8.0")


@pytest.mark.parametrize("args,msg", [
    ((), "required arguments were not provided"),
    (("-r", "lib.json", "-o", "out.tsv"), "required arguments were not provided"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "-c", "many"), "integer value for the number of cores"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "-f", "sideways"), "Could not parse strand_filter option."),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "-t", "40"), "Invalid strictness"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "-t", "x:0.5"), "Invalid length"),
    (("-r", "a.json", "b.json", "-o", "a.tsv", "b.tsv", "-i", "a.fastq", "-t", "40:0.5"), "number of trim options does not match"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "-d", "0,x"), "comma-separated list of GPU ordinals"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "-c"), "requires a value"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "a.fastq", "--index-cache"), "requires a value"),
    (("-r", "lib.json", "-o", "out.tsv", "-i", "reads.sam"), "Unsupported file format: sam"),
    (("-r", "a.json", "b.json", "-o", "a.tsv", "-i", "a.fastq"), "one output path per reference library"),
    (("stray",), "wasn't expected"),
])
def test_argument_errors_exit_like_a_panic(args, msg):
    rc, out = run(*args)
    assert rc == 101 and msg in out, (rc, out)


def test_missing_library_file_fails_before_any_gpu_work(tmp_path):
    fq = tmp_path / "a.fastq"
    fq.write_text("@r\nACGT\n+\nIIII\n")
    rc, out = run("-r", str(tmp_path / "nope.json"), "-o", str(tmp_path / "o.tsv"), "-i", str(fq))
    assert rc == 101 and "Processing as FASTQ file" in out and not (tmp_path / "o.tsv").exists()
    rc, out = run("-r", str(tmp_path / "nope.json"), "-o", str(tmp_path / "o.tsv.gz"), "-i", str(tmp_path / "x.BAM"))
    assert rc == 101 and "Processing as BAM file" in out
