#!/usr/bin/env python3
"""Byte-compile the reference's Python package into sourceless .pyc files (test infrastructure, see oracle/Makefile).
usage: compile_pyref.py <reference python/audioflux dir> <output dir>"""
import os
import py_compile
import sys
import warnings


def main(src, dst):
    warnings.simplefilter("ignore")          # the reference's docstrings contain invalid escape sequences
    n = 0
    for root, _, files in os.walk(src):
        rel = os.path.relpath(root, src)
        for f in files:
            if not f.endswith(".py"):
                continue
            out_dir = os.path.join(dst, rel) if rel != "." else dst
            os.makedirs(out_dir, exist_ok=True)
            py_compile.compile(os.path.join(root, f), cfile=os.path.join(out_dir, f + "c"),
                               dfile=os.path.join("audioflux", rel, f) if rel != "." else os.path.join("audioflux", f),
                               doraise=True)
            n += 1
    print(f"compiled {n} modules into {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
