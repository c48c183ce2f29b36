#!/usr/bin/env python3
"""Regenerate tests/golden/reference_fixtures.json from the reference tree.

The reference (DaymudeLab/EvoSOPS) ships no tests; the only artefacts that pin its behaviour are the
lattice snapshots its authors pasted into the TikZ helper scripts and the genomes in the comparison
scripts.  This script lifts exactly those literals (data, not code) into one JSON fixture:

  misc_scripts/generate_tikz_img_agg.py:1,30,58,86     4 aggregation lattices (13,9) at steps 0,n,n^2,n^3
  misc_scripts/generate_tikz_img_sep.py:1,29,57,85     4 separation lattices (13,9)
  misc_scripts/generate_tikz_img_coat.py:2..296        8 coating lattices (20,5,2)
  misc_scripts/genome_agg_compare.py:1-2               evolved + theory aggregation genomes
  misc_scripts/genome_sep_compare.py:1-2               evolved + theory separation genomes

Run here (the reference is not present on the GPU box):  python tests/golden/make_golden.py
"""
import ast
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")


def literals(path, names):
    """Return every `name = <list literal>` in the file, commented out or not, in file order."""
    text = open(path).read()
    text = "\n".join(re.sub(r"^#\s?", "", ln) for ln in text.splitlines())
    out = []
    for m in re.finditer(r"^\s*(%s)\s*=\s*\[" % "|".join(names), text, flags=re.M):
        start = m.end() - 1
        depth, k = 0, start
        while True:
            ch = text[k]
            if ch == "[":
                depth += 1
            elif ch == "]":
                depth -= 1
                if depth == 0:
                    break
            k += 1
        try:
            out.append((m.group(1), ast.literal_eval(text[start:k + 1])))
        except (SyntaxError, ValueError):
            pass
    return out


def main():
    ms = os.path.join(REF, "misc_scripts")
    fx = {"source": "DaymudeLab/EvoSOPS misc_scripts (see make_golden.py docstring)"}
    agg = [g for _, g in literals(os.path.join(ms, "generate_tikz_img_agg.py"), ["grid"])]
    sep = [g for _, g in literals(os.path.join(ms, "generate_tikz_img_sep.py"), ["grid"])]
    coat = [g for _, g in literals(os.path.join(ms, "generate_tikz_img_coat.py"), ["grid[0-9]"])]
    assert len(agg) == 4 and len(sep) == 4 and len(coat) == 8, (len(agg), len(sep), len(coat))
    for g in agg + sep:
        assert len(g) == 27 and all(len(r) == 27 for r in g)
    for g in coat:
        assert len(g) == 41 and all(len(r) == 41 for r in g)
    fx["agg_13_9_snapshots"] = agg      # steps 0, n, n^2, n^3
    fx["sep_13_9_snapshots"] = sep
    fx["coat_20_5_2_snapshots"] = coat
    ga = dict(literals(os.path.join(ms, "genome_agg_compare.py"), ["g_ga", "g_theory"]))
    gs = dict(literals(os.path.join(ms, "genome_sep_compare.py"), ["g_ga", "g_theory"]))
    fx["agg_genome_evolved"] = ga["g_ga"]
    fx["agg_genome_theory"] = ga["g_theory"]
    fx["sep_genome_evolved"] = gs["g_ga"]
    fx["sep_genome_theory"] = gs["g_theory"]
    with open(OUT, "w") as f:
        json.dump(fx, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
