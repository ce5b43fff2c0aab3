#!/usr/bin/env python
"""Generates tests/golden/reader_golden.json: what the REFERENCE's own FASTA/FASTQ reader
(/root/reference/src/sequence_io.cpp, compiled where it lies with tests/host/seqio_dump_ref.cpp) returns for a
set of tricky files, and what its print_alignment (src/alignment_io.cpp, tests/host/alnio_dump_ref.cpp) prints.  The reference cannot travel to the GPU box, its answers can.

    python tests/golden/make_reader_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/src"

FILES = {
    "a.fa": ">r1 first\nACGT\nacgtNN\n>r2\nGG\n\nTT\n>r3 empty\n>r4\nA\n",
    "crlf.fasta": ">r1\r\nACGT\r\nAC\r\n>r2\r\nTT\r\n",
    "nonl.fna": ">only\nACGTACGT",
    "b.fq": "@q1 x\nACGT\n+\nIIII\n@q2\nGGTT\n+q2\n!!!!\n",
    "c.fastq": "@q1\nAC\n+\nII\n",
    "sniff_fa.txt": ">x\nAAA\nCCC\n",
    "sniff_fq.dat": "@y\nTTT\n+\nIII\n",
    "empty.fa": "",
    "junk.txt": "hello\nworld\n",
    "lead.fa": "\n\n>r1\nAC\n",
    "semi.fa": ";comment\n>r1\nACGT\n",
    "trunc.fq": "@q1\nACGT\n+\n",
    "multi.fnq": "@a\nAC\n+\nII\n@b\nGT\n+\nII\n@c\nTT\n+\nII\n",
    "long.fasta": ">big\n" + "\n".join("ACGTTGCA" * 10 for _ in range(50)) + "\n>second\nTTTT\n",
}


# print_alignment (src/alignment_io.cpp:13-38): "score|q|s|width" lines
ALN_CASES = ["654|ACGT_AC  |AC_TTAC  |80", "-3|   ACGTACGTAC|   AC_TAC__AC|4", "0|||80", "12|A|A|1",
             "7|" + "ACGT" * 50 + "|" + "ACGA" * 50 + "|80", "5|ACGT|ACG|80", "5|ACG|ACGT|3", "-2147483647|  __AA|  TT__|5",
             "9|ACGTAC|ACGTAC|6", "9|ACGTACG|ACGTACG|6"]


def build_ref_printer(outdir):
    exe = os.path.join(outdir, "alnio_dump_ref")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++14", "-I" + REF, os.path.join(ROOT, "tests", "host", "alnio_dump_ref.cpp"),
                    os.path.join(REF, "alignment_io.cpp"), "-o", exe], check=True)
    return exe


def run_printer(exe):
    r = subprocess.run([exe], input="\n".join(ALN_CASES) + "\n", capture_output=True, text=True)
    return {"rc": r.returncode, "stdout": r.stdout}


def build_ref_driver(outdir):
    exe = os.path.join(outdir, "seqio_dump_ref")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++14", "-I" + REF, os.path.join(ROOT, "tests", "host", "seqio_dump_ref.cpp"),
                    os.path.join(REF, "sequence_io.cpp"), "-o", exe], check=True)
    return exe


def run_all(exe, workdir):
    out = {}
    for name, content in FILES.items():
        path = os.path.join(workdir, name)
        with open(path, "w", newline="") as f:
            f.write(content)
        for skip in (0, 2):
            r = subprocess.run([exe, path, str(skip)], capture_output=True, text=True)
            out[f"{name}|skip={skip}"] = {"rc": r.returncode, "stdout": r.stdout}
    return out


def main():
    if not os.path.exists(os.path.join(REF, "sequence_io.cpp")):
        sys.exit("reference sources not found")
    with tempfile.TemporaryDirectory() as td:
        res = run_all(build_ref_driver(td), td)
        prn = run_printer(build_ref_printer(td))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reader_golden.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_reader_golden.py", "files": FILES, "expected": res,
                   "print_alignment_cases": ALN_CASES, "print_alignment": prn}, f, indent=0)
    print("wrote", path, len(res), "cases")


if __name__ == "__main__":
    main()
