#!/usr/bin/env python3
"""One-off importer for the nucleoside alphabet DATA (not code).

Reads the reference's mass tables (assets/masses.tsv, assets/element_masses.tsv under
/root/reference/spectrseqtools) and writes them as one JSON document that
spectrseqtools_b200/masses.py loads.  Values are carried as exact decimal floats (repr round-trips),
so the 4-dp grouping and the 1 mDa integer masses come out identical to the reference
(masses.py:53-88 there).  Run in the build container only; the JSON is committed.
"""
import csv
import json
import pathlib
import sys

ref = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/spectrseqtools/assets")
out = pathlib.Path(__file__).resolve().parents[1] / "spectrseqtools_b200" / "assets" / "alphabet.json"

with open(ref / "masses.tsv", newline="") as fh:
    rows = list(csv.reader(fh, delimiter="\t"))
assert rows[0] == ["nucleoside", "canonical_name", "monoisotopic_mass", "modification_rate"]
nucleosides = [[r[0], r[1], float(r[2]), float(r[3])] for r in rows[1:] if r]

with open(ref / "element_masses.tsv", newline="") as fh:
    erows = list(csv.reader(fh, delimiter="\t"))
assert erows[0] == ["symbol", "mass"]
elements = {r[0]: float(r[1]) for r in erows[1:] if r}

doc = {
    "columns": ["nucleoside", "canonical_name", "monoisotopic_mass", "modification_rate"],
    "nucleosides": nucleosides,
    "elements": elements,
}
lines = ["{", ' "columns": ' + json.dumps(doc["columns"]) + ",", ' "elements": ' + json.dumps(elements) + ",", ' "nucleosides": [']
lines += ["  " + json.dumps(n, ensure_ascii=False) + ("," if i + 1 < len(nucleosides) else "") for i, n in enumerate(nucleosides)]
lines += [" ]", "}"]
out.write_text("\n".join(lines) + "\n")
assert json.loads(out.read_text()) == doc
print(f"wrote {out}: {len(nucleosides)} nucleosides, {len(elements)} elements")
