"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/affine_me.h
declares, fails loudly without a GPU, its geometry agrees with the oracle's tables, and the host logic
(CLI flags, schedule, log format, multi-rank plumbing) behaves like the reference's."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import ame_logs
import oracle_binding as ob
from conftest import load_pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "affine_me.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ame_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    L = pkg.lib()
    decl = _declared_symbols()
    assert len(decl) >= 17
    for s in decl:
        assert hasattr(L, s), s
    assert sorted(pkg.EXPORTS) == decl
    assert L.ame_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = load_pkg()
    with pytest.raises(pkg.AmeError, match="no CPU fallback"):
        pkg.AffineME(416, 240)


def test_product_geometry_matches_oracle_tables():
    pkg = load_pkg()
    for pred, ha, n in ((0, 0, 201), (1, 0, 201), (2, 1, 284), (3, 1, 284)):
        for k in range(n):
            assert pkg.cu_geometry(pred, k) == ob.cu_geometry(ha, k), (pred, k)
        assert pkg.cu_geometry(pred, n)[0] == -1
    assert pkg.num_ctus(1920, 1080) == 135 and pkg.num_ctus(3840, 2160) == 510 and pkg.num_ctus(416, 240) == 8
    assert pkg.num_ctus(7680, 4320) == 2040


def test_product_does_not_reference_oracle():
    """The shipped path must not route through the oracle (no include / link / import)."""
    pdir = os.path.join(ROOT, "vvc-affine-gpu_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".cu", ".h", ".cpp", ".py")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)
    ldd = subprocess.run(["ldd", os.path.join(pdir, "libaffine_me.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


CLI = os.path.join(ROOT, "vvc-affine-gpu_b200", "bin", "affine_b200")


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_flag_surface():
    r = subprocess.run([CLI, "-h"], capture_output=True, text=True)
    assert r.returncode == 1  # main.cpp:75-78
    for flag in ("--DeviceIndex", "-q [ --QP ]", "-f [ --FramesToBeEncoded ]", "--ExtraGradientIter", "-s [ --Resolution ]",
                 "-o [ --OriginalFrames ]", "-r [ --ReferenceFrames ]", "-l [ --CpmvLogFile ]"):
        assert flag in r.stdout, flag
    r = subprocess.run([CLI, "-f", "2", "-s", "416x240"], capture_output=True, text=True)
    assert r.returncode == 1
    assert "[!] ERROR: QP not set." in r.stdout and "[!] ERROR: Input original frames not set." in r.stdout
    assert "Exiting after finding errors in input parameters" in r.stdout
    r = subprocess.run([CLI, "--QP=32", "-f2", "--Resolution", "416x240", "-o", "a", "-r", "b", "--ExtraGradientIter", "2"],
                       capture_output=True, text=True)
    assert "QP=32" in r.stdout and "FramesToBeEncoded=2" in r.stdout
    assert "Using a total of 7 iterations for 2 CPs and 6 iterations for 3 CPs." in r.stdout
    r = subprocess.run([CLI, "--bogus"], capture_output=True, text=True)
    assert r.returncode != 0


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
@pytest.mark.parametrize("qp", [22, 27, 32, 37])
def test_product_schedule_matches_oracle(qp):
    """The PRODUCT's reference-list and lambda schedule (host/schedule.cpp, compiled into affine_b200) for POC 1..64,
    long-term references included, against the oracle's restatement of main.cpp:591-707 and
    main_aux_functions.h:1473-1497.  The CLI prints its plan (testReferences, main_aux_functions.h:1499-1545; that
    loop stops one frame short, hence -f 65) before it touches the GPU, so this runs without one: the exit code is
    not looked at (without a device the CLI stops at the pinned allocation, with one at the missing input files)."""
    import re
    r = subprocess.run([CLI, "-f", "65", "-s", "416x240", "-q", str(qp), "-o", "/nonexistent/o.csv", "-r", "/nonexistent/r.csv"],
                       capture_output=True, text=True)
    rows = re.findall(r"^POC +(\d+) +QP (\d+) motionLambda ([0-9.]+) : \[L0([ 0-9-]*)\]$", r.stdout, re.M)
    assert [int(x[0]) for x in rows] == list(range(1, 65)), r.stdout[-1500:]
    lists = ob.ref_lists(64)
    for poc_s, q_s, lam_s, refs_s in rows:
        poc = int(poc_s)
        assert [int(v) for v in refs_s.split()] == lists[poc - 1], poc
        assert int(q_s) == ob.delta_qp(qp, poc), poc
        assert lam_s == "%f" % ob.lambda_for(qp, poc), poc
    # the long-term branch was really exercised: POC 64 keeps POC%8==0 frames far behind it
    assert lists[63] != [63, 62, 61, 60] and any(v % 8 == 0 and v < 56 for v in lists[63])


def test_bench_schedule_helpers_match_oracle():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.ref_lists(64) == ob.ref_lists(64)
    for qp in (22, 27, 32, 37):
        for poc in range(1, 40):
            assert bench.lambda_for(qp, poc) == ob.lambda_for(qp, poc)
    assert bench.OPS_PER_PASS == pytest.approx(40.5e9, rel=0.01)


def test_bench_shard_plan_covers_the_batch_once():
    """configs[4]: the 4096-frame batch is dealt to the ranks in blocks of 8 frames; every block exactly once."""
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            plan = bench.step_plan(64, 64, True, rank, world)
            frames = sum(f1 - f0 for seq in plan for (f0, f1) in seq)
            assert frames == 4096 // world
            assert all(len(seq) <= 8 and all(f1 - f0 == bench.SHARD_BLOCK and f0 % 8 == 0 for f0, f1 in seq) for seq in plan)
            seen += [(i, j) for i, seq in enumerate(plan) for j in range(len(seq))]
        assert len(seen) == 512
    assert bench.step_plan(64, 1, False, 3, 8) == [[(0, 64)]]
    # S of SURVEY 8(d), recomputed from the geometry tables
    assert bench.in_frame_samples(1920, 1080) == bench.S_1080P
    assert bench.in_frame_samples(3840, 2160) == bench.S_BY_SIZE[(3840, 2160)]


def test_log_reader_roundtrip(tmp_path):
    """ame_logs.read_logs inverts the reference's log layout (group files shared by several HA groups)."""
    W, H, n = 416, 240, 2
    nct = ame_logs.num_ctus(W, H)
    rng = np.random.default_rng(0)
    truth = {}
    for (poc, r) in ame_logs.pass_list(n):
        truth[(poc, r)] = []
        for pred in range(4):
            st, total = ame_logs.strides(pred)
            c = rng.integers(0, 1 << 20, nct * total)
            m = np.zeros(nct * total, ame_logs.CPMV_DTYPE)
            for f in ("LTx", "LTy", "RTx", "RTy", "LBx", "LBy"):
                m[f] = rng.integers(-1000, 1000, nct * total)
            truth[(poc, r)].append((c, m))
    prefix = str(tmp_path / "log")
    for pred in range(4):
        st, total = ame_logs.strides(pred)
        opened = set()
        for (poc, r) in ame_logs.pass_list(n):
            c, m = truth[(poc, r)][pred]
            for g, (w, h, cnt) in enumerate(ame_logs.groups(pred)):
                path = "%s%s%dx%d.csv" % (prefix, ame_logs.PRED_TAGS[pred], w, h)
                with open(path, "a" if path in opened else "w") as f:
                    if path not in opened:
                        f.write(ame_logs.HEADER + "\n")
                        opened.add(path)
                    for ctu in range(nct):
                        for i in range(cnt):
                            k = ctu * total + st[g] + i
                            f.write("%d,0,%d,%d,%d,0,0,%d,%d,%d,%d,%d,%d,%d\n" % (poc, r, ctu, i, c[k], m["LTx"][k], m["LTy"][k],
                                                                                     m["RTx"][k], m["RTy"][k], m["LBx"][k], m["LBy"][k]))
    got = ame_logs.read_logs(prefix, W, H, n)
    for key in truth:
        for pred in range(4):
            assert (got[key][pred][0] == truth[key][pred][0]).all()
            assert (got[key][pred][1] == truth[key][pred][1]).all()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # the N>1 plumbing of bench.py: barrier, max-over-ranks of the device time, whole-job frames
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frames = torch.tensor([64.0])
    dist.all_reduce(frames)
    q.put((rank, float(t.item()), float(frames.item())))
    dist.destroy_process_group()


def test_multi_rank_plumbing_gloo():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
    assert [r[1] for r in res] == [11.0, 11.0]      # max over ranks
    assert [r[2] for r in res] == [128.0, 128.0]    # weak scaling: every rank contributes its own 64 frames


@pytest.mark.parametrize("size", ["416x240", "424x236", "1000x600", "1920x1080", "3840x2160"])
def test_tiled_plane_layout_reader_matches_writer(size, tmp_path_factory):
    """Host-side check of the tiled layout of the pre-filtered planes (ame_device.h: tile_record): unique indices inside
    the allocation, and every nine-row window (first row through tile_record, then constant steps) lands on the records
    the phase kernel's addressing stores for those rows, halo rows included (tests/cpp/tile_layout_check.cpp)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("needs g++ and the CUDA headers")
    exe = os.path.join(str(tmp_path_factory.getbasetemp()), "tile_layout_check")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I/usr/local/cuda/include", "-o", exe,
                               os.path.join(ROOT, "tests", "cpp", "tile_layout_check.cpp")])
    w, h = size.split("x")
    r = subprocess.run([exe, w, h], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok ")
