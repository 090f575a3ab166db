"""Diagnostic: doctest_1d fixture on the explicit-rows path; prints where G / Q differ from the golden fixture."""
import os, sys
os.environ["SAFEOPT_B200_GRID_FAST_PATH"] = "0"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from conftest import golden_problem, load_golden, unpack_mask, golden_lipschitz
import safeopt_b200 as sb
for name in sys.argv[1:] or ["doctest_1d"]:
    g = load_golden(name)
    n_rows = int(g["n_rows"])
    gps, grid, fmin = golden_problem(g, "gpu")
    opt = sb.SafeOpt(gps if len(gps) > 1 else gps[0], grid, fmin if len(fmin) > 1 else fmin[0], beta=float(g["beta"]),
                     threshold=float(g["threshold"]), lipschitz=golden_lipschitz(g))
    if bool(g["full_sets"]):
        opt.update_confidence_intervals(); opt.compute_sets(full_sets=True); x = opt.get_new_query_point()
    else:
        x = opt.optimize()
    G_ref = unpack_mask(g["G"], n_rows)
    print(name, "maxdQ", np.abs(opt.Q - g["Q"]).max(), "S eq", np.array_equal(opt.S, unpack_mask(g["S"], n_rows)),
          "M eq", np.array_equal(opt.M, unpack_mask(g["M"], n_rows)), "G ours", np.flatnonzero(opt.G)[:20], "G ref", np.flatnonzero(G_ref)[:20],
          "nG", opt.G.sum(), G_ref.sum(), "row", opt.last_query_row, int(g["row_next"]))
