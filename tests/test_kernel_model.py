"""CPU: the Python model of the DEVICE algorithms (row masks, first-visit replay, path enumeration)
against the reference-made golden vectors.  Proves the restatement the kernels implement, not the kernels."""
import helpers as Hh
import kernel_model as KM
from oracle import oracle_c as OC
from oracle import oracle_py as OP


def test_model_matches_reference_on_small_alphabets():
    cases = Hh.load_json("explain_small.json.gz")
    tables = {}
    modes = {"free": 0, "memo": 0, "exact": 0}
    for c in cases[::2]:
        w = c["weights"]
        tab = tables.setdefault(tuple(w), OC.build_bit_table(w, max(w) * 35, 32))
        rows = [OP.Row(m, im, rt) for m, im, rt in zip(w, c["is_mod"], c["rates"])]
        ind = OP.individual_budgets(rows, c["max_len"])
        target, thr = OP.integerise(c["mass"], c["threshold"], 1e-3, c["tolerance"])
        mm = KM.INF if c["max_modifications"] is None else c["max_modifications"]
        for memo, tag in ((True, "memo"), (False, "nomemo")):
            gold = c[f"solutions_{tag}"]
            if gold == "NotImplementedError":
                try:
                    KM.explain(tab, w, c["is_mod"], ind, target, thr, mm, memo)
                    raise AssertionError("expected NotImplementedError")
                except NotImplementedError:
                    continue
            sols, zero, mode = KM.explain(tab, w, c["is_mod"], ind, target, thr, mm, memo)
            modes[mode] += 1
            assert sorted(tuple(w[r] for r in s) for s in sols) == sorted(tuple(s) for s in gold if s), (c, tag)
            assert zero == any(len(s) == 0 for s in gold)
    assert min(modes.values()) > 100, modes
