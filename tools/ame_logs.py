"""Reader for the reference's per-CU decision logs (the `-l <prefix>` CSV files,
main_aux_functions.h:387-525) -> result arrays indexed like the kernels' output
buffers (ctu*201 + stride + idx / ctu*284 + stride + idx).

Used by tests and tools to compare logs written by the reference binary, by the
oracle and by this repo's CLI.  The group tables below restate SURVEY.md
Appendix B (file naming and group order), independently of the product code.
"""
import os

import numpy as np

CPMV_DTYPE = np.dtype([("nCPs", "<i4"), ("LTx", "<i4"), ("LTy", "<i4"), ("RTx", "<i4"),
                       ("RTy", "<i4"), ("LBx", "<i4"), ("LBy", "<i4")])
PRED_TAGS = ("_FULL_2CPs_", "_FULL_3CPs_", "_HALF_2CPs_", "_HALF_3CPs_")
# (w, h, nCUs) in result order
ALIGNED = [(128, 128, 1), (128, 64, 2), (64, 128, 2), (64, 64, 4), (64, 32, 8), (32, 64, 8), (32, 32, 16),
           (64, 16, 16), (16, 64, 16), (32, 16, 32), (16, 32, 32), (16, 16, 64)]
HALF = [(64, 32, 4), (32, 64, 4), (64, 16, 8), (64, 16, 4), (16, 64, 8), (16, 64, 4), (32, 32, 8), (32, 32, 8),
        (32, 16, 16), (32, 16, 8), (32, 16, 16), (16, 32, 16), (16, 32, 8), (16, 32, 16), (16, 16, 32), (16, 16, 32),
        (16, 16, 16), (16, 16, 16), (32, 32, 4), (32, 16, 8), (32, 16, 4), (16, 32, 8), (16, 32, 4), (16, 16, 32)]
HEADER = "POC,List,Ref,CTU,idx,X,Y,Cost,LT_X,LT_Y,RT_X,RT_Y,LB_X,LB_Y"


def groups(pred):
    return ALIGNED if pred < 2 else HALF


def strides(pred):
    out, s = [], 0
    for (_, _, n) in groups(pred):
        out.append(s)
        s += n
    return out, s


def num_ctus(W, H):
    return ((W + 127) // 128) * ((H + 127) // 128)


def pass_list(n_frames):
    """(poc, refIdx) pairs in the order the reference processes them."""
    return [(poc, r) for poc in range(1, n_frames + 1) for r in range(min(4, poc))]


def read_logs(prefix, W, H, n_frames):
    """Returns {(poc, refIdx): [(costs, cpmvs) for pred in 0..3]} and checks the
    constant columns (List, CTU, idx, X, Y) of every row."""
    nct = num_ctus(W, H)
    cols = (W + 127) // 128
    passes = pass_list(n_frames)
    res = {p: [None] * 4 for p in passes}
    for pred in range(4):
        grp = groups(pred)
        st, total = strides(pred)
        for p in passes:
            res[p][pred] = (np.zeros(nct * total, np.int64), np.zeros(nct * total, CPMV_DTYPE))
        names = []
        for (w, h, _) in grp:
            nm = "%dx%d" % (w, h)
            if nm not in names:
                names.append(nm)
        for nm in names:
            path = prefix + PRED_TAGS[pred] + nm + ".csv"
            with open(path) as f:
                head = f.readline().strip()
                assert head == HEADER, (path, head)
                data = np.loadtxt(f, delimiter=",", dtype=np.int64, ndmin=2)
            gs = [g for g, (w, h, _) in enumerate(grp) if "%dx%d" % (w, h) == nm]
            rows_per_pass = sum(grp[g][2] for g in gs) * nct
            assert data.shape == (rows_per_pass * len(passes), 14), (path, data.shape)
            row = 0
            for (poc, r) in passes:
                costs, cp = res[(poc, r)][pred]
                for g in gs:
                    w, h, n = grp[g]
                    blk = data[row:row + n * nct]
                    row += n * nct
                    ctu = np.repeat(np.arange(nct), n)
                    idx = np.tile(np.arange(n), nct)
                    assert (blk[:, 0] == poc).all() and (blk[:, 1] == 0).all() and (blk[:, 2] == r).all(), path
                    assert (blk[:, 3] == ctu).all() and (blk[:, 4] == idx).all(), path
                    dst = ctu * total + st[g] + idx
                    costs[dst] = blk[:, 7]
                    for k, fld in enumerate(("LTx", "LTy", "RTx", "RTy", "LBx", "LBy")):
                        cp[fld][dst] = blk[:, 8 + k]
    return res


def log_files(prefix):
    d = os.path.dirname(prefix) or "."
    b = os.path.basename(prefix)
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.startswith(b + "_") and f.endswith(".csv"))
