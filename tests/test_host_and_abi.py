"""CPU: host-side logic of the product and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.fft as sfft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_library_exports_every_declared_symbol():
    from fcd_b200 import _native, build
    lib_path = build.build()
    header = open(os.path.join(ROOT, "include", "fcd_b200.h")).read()
    declared = set(re.findall(r"\b(fcd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    # the product library is CUDA code for sm_100a
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_wavenumber_helpers_match_scipy():
    from fcd_b200 import engine as e
    for n in (64, 256, 1024):
        for cal in (1, 0.37, 2.0e-4):
            assert np.array_equal(e.wavenumber(n, cal), sfft.fftfreq(n, cal / (2 * np.pi)))
            assert np.array_equal(e.wavenumber(n, cal, True), sfft.fftshift(sfft.fftfreq(n, cal / (2 * np.pi))))
    from oracle import fcd_oracle as o
    pk = [np.array([142, 144]), np.array([112, 142])]
    assert np.array_equal(e.pixel_to_wavenumber((256, 256), pk), o.pixel_to_wavenumber((256, 256), pk))
    assert np.array_equal(e.pixel_to_wavenumber((256, 512), pk[0], 0.5), o.pixel_to_wavenumber((256, 512), pk[0], 0.5))
    assert e.calibration_from_peaks((256, 256), pk, 256 / 30) == 1.0


def test_height_resolution_like_reference(golden):
    from fcd_b200 import engine as e
    layers = golden["layers.example"].tolist()
    assert e.height_from_layers(layers) == float(golden["layers.height"])
    assert e.resolve_height() == 1 and e.resolve_height(height=0.3) == 0.3
    assert e.resolve_height(layers=layers) == float(golden["layers.height"])
    with pytest.raises(Warning):
        e.resolve_height(layers=layers, height=1.0)


def test_shard_ranges_partition_the_frames():
    from fcd_b200 import engine as e
    for n in (0, 1, 7, 1000, 20000):
        for w in (1, 2, 4, 8):
            r = [e.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        e.shard_range(10, 2, 2)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pyfcd.fcd import fcd
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fcd.compute_height_map(np.zeros((64, 64), np.float32), np.zeros((64, 64), np.float32), 1.0)


_WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "trapped-modes-ltg_b200"))
from fcd_b200 import engine as e
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n = 11
full = torch.arange(n * 4 * 5, dtype=torch.float32).reshape(n, 4, 5)
a, b = e.shard_range(n, dist.get_rank(), 2)
out = e.gather_height_maps(full[a:b].clone(), n, dst=0, chunk_frames=2)
if dist.get_rank() == 0:
    assert torch.equal(out, full)
    print("GATHER_OK")
else:
    assert out is None
dist.barrier(); dist.destroy_process_group()
'''


def test_sharded_gather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


_RING_WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "trapped-modes-ltg_b200"))
from fcd_b200.peer_ring import PeerRing
rank, world = int(sys.argv[3]), 3
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=world)
import time
chunk, nchunks = 5, [0, 90, 70]                    # rank 1 pushes 90 chunks, rank 2 pushes 70 (the last ones ragged)
seen = {}
def consume(src, k, slot, cnt, t0, t1):
    seen[(src, k)] = (cnt, float(slot[:cnt].sum()), t0, t1)
for slots in (2, 4):
    seen.clear()
    ring = PeerRing((chunk, 3, 4), nchunks, root=0, slots=slots, device="cpu", consume=consume)
    if rank != 0:
        for k in range(nchunks[rank]):
            cnt = chunk if k + 1 < nchunks[rank] else 2
            buf = torch.full((chunk, 3, 4), float(100 * rank + k))
            time.sleep(0.002)
            ring.push(buf, cnt, tag0=k, tag1=rank)
    else:
        time.sleep(0.1)        # the root finishes its own work mid-stream and closes while its consumer is still busy
    ring.close()
    if rank == 0:
        assert len(seen) == 160, len(seen)
        for (src, k), (cnt, total, t0, t1) in seen.items():
            assert cnt == (chunk if k + 1 < nchunks[src] else 2) and (t0, t1) == (k, src)
            assert total == cnt * 12 * (100 * src + k)
        print("RING_OK", slots)
dist.barrier(); dist.destroy_process_group()
'''


def test_peer_ring_control_protocol_three_ranks_gloo(tmp_path):
    """The streamed gather's ring (fcd_b200/peer_ring.py) in its CPU mode: slot-filled / slot-free messages with
    2 and 4 slots, ragged last chunks, two senders, the root closing mid-stream (a hang on 8 GPUs came from exactly
    that) -- the data plane is shared memory here, NVLink on the box."""
    script = tmp_path / "ring_worker.py"
    script.write_text(_RING_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(3)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "RING_OK 2" in outs[0] and "RING_OK 4" in outs[0]


def test_load_image_gray_and_colour(tmp_path):
    """analyze.load_image (pydata/analyze.py:26-40: skimage.io.imread(as_gray=True).astype(float32)): 2-D files keep
    their values, colour files take scikit-image's luminance (restated; scikit-image is not installed here)."""
    import cv2
    from pydata.analyze import analyze
    rng = np.random.default_rng(1)
    g16 = rng.integers(0, 1024, (12, 9), dtype=np.uint16)
    cv2.imwrite(str(tmp_path / "g.tif"), g16)
    out = analyze.load_image(str(tmp_path / "g.tif"))
    assert out.dtype == np.float32 and np.array_equal(out, g16.astype(np.float32))
    rgb = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)                      # R, G, B
    cv2.imwrite(str(tmp_path / "c.png"), rgb[..., ::-1])                       # OpenCV writes B, G, R
    want = (rgb.astype(np.float64) / 255.0) @ np.array([0.2125, 0.7154, 0.0721])
    got = analyze.load_image(str(tmp_path / "c.png"))
    assert got.dtype == np.float32 and np.allclose(got, want, rtol=0, atol=1e-7)
    rgba = np.concatenate([rgb, rng.integers(0, 256, (7, 5, 1), dtype=np.uint8)], axis=2)
    cv2.imwrite(str(tmp_path / "a.png"), rgba[..., [2, 1, 0, 3]])
    a = rgba[..., 3:4] / 255.0
    want = ((1 - a) + a * rgba[..., :3] / 255.0) @ np.array([0.2125, 0.7154, 0.0721])
    assert np.allclose(analyze.load_image(str(tmp_path / "a.png")), want, rtol=0, atol=1e-7)
    with pytest.raises(FileNotFoundError):
        analyze.load_image(str(tmp_path / "missing.tif"))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on CPU: exactly one JSON line on stdout with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--size", "256",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_mask_bytes_reinterprets_instead_of_copying():
    """Masks reach the kernels as one byte per pixel, zero = keep.  bool / uint8 masks must not be copied (a dtype
    copy of a batch of 2048^2 masks costs more than K1); other dtypes become 0 / 1 bytes of their truth value."""
    import torch
    from fcd_b200 import engine as e
    b = torch.zeros((3, 8, 8), dtype=torch.bool); b[1, 2:5, 3] = True
    v = e._mask_bytes(b, torch.device("cpu"))
    assert v.dtype == torch.uint8 and v.data_ptr() == b.data_ptr() and v.tolist() == b.to(torch.uint8).tolist()
    u = torch.tensor([[0, 1], [7, 0]], dtype=torch.uint8)
    assert e._mask_bytes(u, torch.device("cpu")).data_ptr() == u.data_ptr()
    w = e._mask_bytes(np.array([[0.0, -2.5], [256.0, 0.0]]), torch.device("cpu"))      # 256 must not wrap to 0
    assert w.dtype == torch.uint8 and w.tolist() == [[0, 1], [1, 0]]
    nb = e._mask_bytes(np.array([[True, False]]), torch.device("cpu"))
    assert nb.dtype == torch.uint8 and nb.tolist() == [[1, 0]]
    t = torch.ones((4, 6), dtype=torch.bool)[:, ::2]                                    # non-contiguous view
    assert e._mask_bytes(t, torch.device("cpu")).is_contiguous()


def test_launcher_tile_coordinates_model():
    """csrc/fcd_launch.cuh keeps (bx, by) = (tile % gx, tile / gx) incrementally (no division per tile).  Model of that
    loop for both tile orders: every block must visit exactly its tiles with the right coordinates and hand the right
    next-tile link to the kernels that prefetch."""
    def block_tiles(blk, grid, gx, ntiles, blocked):
        t0, t1, step = blk, ntiles, grid
        if blocked:
            t0, t1, step = blk * ntiles // grid, (blk + 1) * ntiles // grid, 1
        if t0 >= t1:
            return []
        bx, by = t0 % gx, t0 // gx
        sx, sy = (1, 0) if blocked else (step % gx, step // gx)
        seen, tile = [], t0
        while tile < t1:
            nbx, nby = bx + sx, by + sy
            if nbx >= gx:
                nbx -= gx; nby += 1
            seen.append((tile, bx, by, nbx, nby, tile + step < t1, tile == t0))
            bx += sx; by += sy
            if bx >= gx:
                bx -= gx; by += 1
            tile += step
        return seen

    for gx, gy, grid in [(1, 1, 1), (1, 7, 3), (5, 1, 2), (128, 512, 148), (3, 1025, 148), (1000, 2, 296), (7, 11, 77), (2, 3, 6)]:
        ntiles = gx * gy
        grid = min(grid, ntiles)
        for blocked in (False, True):
            visited = []
            for blk in range(grid):
                tiles = block_tiles(blk, grid, gx, ntiles, blocked)
                step = 1 if blocked else grid
                for i, (tile, bx, by, nbx, nby, has_next, first) in enumerate(tiles):
                    assert (bx, by) == (tile % gx, tile // gx)
                    assert first == (i == 0) and has_next == (i + 1 < len(tiles))
                    if has_next:
                        assert (nbx, nby) == ((tile + step) % gx, (tile + step) // gx)
                    visited.append(tile)
            assert sorted(visited) == list(range(ntiles))


def test_folder_mask_preview_dialogue(tmp_path, monkeypatch):
    """analyze.folder(show_mask=True) (pydata/analyze.py:193-217): the eleventh frame (or the last one) is previewed,
    "n" asks for a new `smoothed`, a non-integer keeps the previous one, "Y" accepts; without `smoothed` it is a
    ValueError.  matplotlib, the image decode and the device mask are stubbed: this is the dialogue only."""
    import types
    from pydata.analyze import analyze
    events = []
    plt = types.ModuleType("matplotlib.pyplot")
    plt.pause = lambda s: events.append(("pause", s))
    plt.close = lambda what: events.append(("close", what))
    mpl = types.ModuleType("matplotlib"); mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    monkeypatch.setattr(analyze, "load_image", classmethod(lambda cls, path: events.append(("load", os.path.basename(path))) or np.zeros((4, 4), np.float32)))
    monkeypatch.setattr(analyze, "mask", classmethod(lambda cls, image, smoothed=14, show_mask=False, find_center=False:
                                                     events.append(("mask", smoothed, show_mask))))
    answers = iter(["n", "21", "maybe", "n", "x1", "Y"])
    names = [f"f{i:02d}.tif" for i in range(14)]
    got = analyze._preview_mask(str(tmp_path), names, 15, ask=lambda prompt: next(answers))
    assert got == 21
    assert events[0] == ("load", "f10.tif")
    assert [e[1] for e in events if e[0] == "mask"] == [15, 21, 21, 21] and all(e[2] for e in events if e[0] == "mask")
    assert events.count(("pause", 8)) == 4 and events.count(("close", "all")) == 4
    events.clear()
    assert analyze._preview_mask(str(tmp_path), names[:3], 9, ask=lambda prompt: "Y") == 9 and events[0] == ("load", "f02.tif")
    with pytest.raises(ValueError):
        analyze.folder("ref.tif", str(tmp_path), None, 1.0, smoothed=None, show_mask=True)
