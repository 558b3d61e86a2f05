"""Edge-case golden vectors for the PQ / DVPQ evaluators, produced by the REAL reference code (authoring container only):
eval/eval_dvpq.py::vpq_eval and ldmseg/evaluations/cityscapes_pap_eval.py::CityscapesPanopticEvaluator on the inputs of
tests/synth.py::edge_cases_vpq / edge_cases_city.

    python tests/golden/make_golden_edge_cases.py      -> tests/golden/edge_cases.json
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_stubs  # noqa: E402
from synth import edge_cases_city, edge_cases_vpq  # noqa: E402

ref_stubs.install()
sys.argv = [sys.argv[0]]  # eval/eval_dvpq.py parses argv at import


def main():
    dv = ref_stubs.load_by_path("ref_eval_dvpq", "eval/eval_dvpq.py")
    ce = ref_stubs.load_by_path("ref_city_eval", "ldmseg/evaluations/cityscapes_pap_eval.py")
    out = {"vpq": {}, "city": {}}
    for name, (pred, gt) in edge_cases_vpq().items():
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            r = dv.vpq_eval([pred.astype(np.int64), gt.astype(np.int64)])  # int64: the non-overflowing semantics
        out["vpq"][name] = [x.tolist() for x in r]
    for name, (pred, gt) in edge_cases_city().items():
        ev = ce.CityscapesPanopticEvaluator(thing_ids={11, 12, 13, 14, 15, 16, 17, 18})
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            ev.add_image(pred.copy(), gt.copy())
            res = ev.evaluate()
        out["city"][name] = {"tp": int(ev.TP), "fp": int(ev.FP), "fn": int(ev.FN), "iou_sum": float(ev.iou_sum),
                             "pq": float(res["pq"]), "sq": float(res["sq"]), "rq": float(res["rq"]),
                             "per_class": {str(c): {k: (float(v) if isinstance(v, float) else int(v))
                                                    for k, v in m.items()} for c, m in res["per_class"].items()}}
    with open(os.path.join(HERE, "edge_cases.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out["vpq"].items():
        print("vpq", k, "tp", sum(v[1]), "fn", sum(v[2]), "fp", sum(v[3]), "iou", sum(v[0]))
    for k, v in out["city"].items():
        print("city", k, {kk: v[kk] for kk in ("tp", "fp", "fn", "iou_sum")})


if __name__ == "__main__":
    main()
