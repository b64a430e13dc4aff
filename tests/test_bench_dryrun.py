"""Control flow of bench.py's GPU arm without a GPU: the CUDA context, torch.cuda and torch.distributed are replaced by
stand-ins, so that every branch (orbit / every-pair, T,Q,U / TT, one rank / a rank of several, e2e, gather) builds its
JSON line.  Guards against a NameError or a wrong attribute in a script that is otherwise only run on the GPU box."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeEvent:
    def __init__(self, enable_timing=True):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 12.5


class FakeContext:
    """the methods of capi.Context that bench.py and multigpu.py call; buffers are plain host memory"""

    def __init__(self, device=0, stream=None):
        self.stream_handle = None
        self.launches = 0
        self._bufs = {}
        self.npix = 0
        self.variant = 0

    def set_stream(self, s):
        self.stream_handle = s

    def synchronize(self):
        pass

    def set_pixels(self, nside, good=None):
        self.npix = 12 * nside * nside if good is None else len(good)

    def set_kernel_variant(self, v):
        self.variant = v

    def set_host_expand(self, threads):
        self.host_expand = threads

    def measure_fp64_peak(self):
        return 37.0

    def device_malloc(self, nbytes):
        a = np.empty(max(nbytes // 8, 1))
        self._bufs[a.ctypes.data] = a
        return a.ctypes.data

    def device_free(self, ptr):
        del self._bufs[ptr]

    def _launch(self, *a, **k):
        self.launches += 1

    legendre_series = legendre_series_orbit = legendre_series_orbit_sharded = tqu = tqu_orbit = tqu_orbit_sharded = _launch
    cl_to_cmatrix = cl_to_cmatrix_pol = cl_to_cmatrix_dev = cl_to_cmatrix_pol_dev = tqu_orbit_assemble = tqu_scatter_block = tqu_orbit_scatter_inbox = tqu_batched_slab = _launch

    def orbit_strips_to_host(self, shard, host, threads=0, direct_mask=0):
        self.to_host_calls = getattr(self, "to_host_calls", 0) + 1

    def ipc_export(self, ptr):
        return b"handle-of-%d" % ptr

    def ipc_open(self, handle):
        return 0x7000000

    def ipc_close(self, ptr):
        pass

    def copy_on_device(self, dst, src, nbytes):
        pass

    def close(self):
        pass


@pytest.fixture
def fake_gpu(monkeypatch):
    import cosmopp_b200
    from cosmopp_b200 import multigpu
    import bench
    monkeypatch.setattr(cosmopp_b200, "Context", FakeContext)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (170 << 30, 180 << 30))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    real_empty, real_tensor = torch.empty, torch.tensor

    def strip(kw):
        kw.pop("device", None)
        kw.pop("pin_memory", None)
        return kw

    monkeypatch.setattr(torch, "empty", lambda *a, **kw: real_empty(*a, **strip(kw)))
    monkeypatch.setattr(torch, "tensor", lambda *a, **kw: real_tensor(*a, **strip(kw)))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    # device buffers as views of the fake context's host arrays
    monkeypatch.setattr(multigpu.DeviceBuffer, "tensor", lambda self: torch.from_numpy(self.ctx._bufs[self.ptr])[:self.n])
    import torch.distributed as dist
    for name in ("init_process_group", "barrier", "destroy_process_group"):
        monkeypatch.setattr(dist, name, lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    monkeypatch.setattr(dist, "broadcast", lambda t, src=0: None)
    monkeypatch.setattr(dist, "all_to_all_single", lambda out, inp, out_splits=None, in_splits=None: None)
    monkeypatch.setattr(dist, "all_gather_object", lambda lst, obj: [lst.__setitem__(i, obj) for i in range(len(lst))])
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    from cosmopp_b200 import capi
    monkeypatch.setattr(capi, "host_register", lambda a: None)
    monkeypatch.setattr(capi, "host_unregister", lambda a: None)
    monkeypatch.setattr(bench, "spot_check", lambda *a, **k: 3e-14)
    import cosmopp_b200.likelihood as lk

    class FakeLike:
        def __init__(self, *a, **k):
            pass

        def calculate(self, t):
            return 1234.5, 1000.0, 234.5

        def close(self):
            pass
    monkeypatch.setattr(lk, "Likelihood", FakeLike)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []})
    monkeypatch.setattr(bench, "reference_sample", lambda *a, **k: {"value": 1.0, "unit": bench.UNIT, "cores": 1, "kind": "reference",
                                                                    "sample": "stub", "wall_s": 1.0})
    return bench


def _run(bench, capsys, monkeypatch, argv, rank=0, world=1):
    monkeypatch.setenv("RANK", str(rank))
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    out = capsys.readouterr().out.strip()
    return json.loads(out.splitlines()[-1]) if out else None


REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


@pytest.mark.parametrize("workload,extra,orbit,launches_per_step", [
    ("tqu_nside32_lmax96", [], True, 1), ("tqu_nside32_lmax96", ["--no-orbit"], False, 1), ("tqu_nside32_lmax96", ["--host-expand", "4"], True, 1),
    ("tt_nside32_lmax96", [], True, 1), ("tt_nside16_lmax47", ["--no-orbit"], False, 1),
    ("tqu_nside16_lmax47_masked", [], False, 1)])
def test_single_rank_line(fake_gpu, capsys, monkeypatch, workload, extra, orbit, launches_per_step):
    line = _run(fake_gpu, capsys, monkeypatch, ["--workload", workload, "--steps", "2", "--warmup", "3"] + extra)
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["gpu_launches"] == 2 * launches_per_step
    assert ("symmetry orbits" in line["path"]["method"]) == orbit
    r = line["roofline"]
    assert r["bound"] == "fp64" and 0 < r["frac"] and r["unit"] == "TFLOP/s"
    if orbit:
        assert r["evaluated_pixel_pairs"] < r["stored_pixel_pairs_all_ranks"] and r["note"]
    else:
        assert r["evaluated_pixel_pairs"] == r["stored_pixel_pairs_all_ranks"]
    e = line["e2e"]
    assert e["value"] > 0 and 0 < e["d2h_bytes_per_step"] <= line["config"]["packed_bytes"] and e["h2d_bytes_per_step"] > 0
    assert (line["parity_max_err"] == 3e-14) == (orbit and workload.startswith("tqu"))
    if line["config"]["packed_bytes"] <= (2 << 30):
        assert line["e2e_device_consumer"]["d2h_bytes_per_step"] == 16 and line["e2e_device_consumer"]["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "reference"
    ref = fake_gpu.config_dict(workload, *fake_gpu.workload_geometry(workload)[:3], line["config"]["npix"], 1)
    assert ref == line["config"]                         # the reference arm describes the same workload


@pytest.mark.parametrize("extra", [[], ["--exchange", "nccl"], ["--no-orbit"], ["--gather"], ["--no-orbit", "--gather"], ["--direct-mask", "0"]])
def test_rank_of_several(fake_gpu, capsys, monkeypatch, extra):
    argv = ["--workload", "tqu_nside32_lmax96", "--gpus", "4", "--steps", "2", "--warmup", "3"] + extra
    assert _run(fake_gpu, capsys, monkeypatch, argv, rank=3, world=4) is None          # only rank 0 prints
    line = _run(fake_gpu, capsys, monkeypatch, argv, rank=0, world=4)
    assert line["n_gpus"] == 4 and line["cpu_baseline"] is None
    assert line["e2e"]["d2h_bytes_per_step"] > 0
    orbit = "--no-orbit" not in extra
    if orbit:
        # orbit shards: exchange timed, whole matrix gathered by default, one shared host matrix every rank writes its columns of
        assert line["exchange"]["ms"] > 0 and line["exchange"]["bytes_sent_this_rank"] > 0 and line["parity_max_err"] == 3e-14
        assert line["gather"]["ms"] > 0 and line["e2e"]["host_matrix_max_abs_diff_vs_device"] is not None
        # N ranks ship at most the matrix once (here: only the last-face columns)
        assert line["e2e"]["d2h_bytes_per_step"] * 4 <= 1.05 * line["config"]["packed_bytes"]
        assert line["exchange"]["mode"] == ("nccl" if "nccl" in extra else "pull")
    if "--gather" in extra:
        assert line["gather"]["ms"] > 0 and line["gather"]["bytes_per_gpu_in"] > 0
    assert ("orbit-closed" in line["path"]["sharding"]) == ("--no-orbit" not in extra)


def test_batched_workload_line(fake_gpu, capsys, monkeypatch):
    """BASELINE configs[3] under the driver's command line: both bounds of the ridge-point kernel in the line"""
    import cosmopp_b200.capi as capi
    monkeypatch.setattr(capi, "slab_doubles", lambda dim: 16 * capi.packed_size(dim))
    monkeypatch.setitem(fake_gpu.WORKLOADS, "batched_x1024_tqu_nside16_lmax47", ("batched", 4, 8, 40))       # small stand-in shape
    monkeypatch.setattr(fake_gpu, "run_batched_arm", fake_gpu.run_batched_arm)
    line = _run(fake_gpu, capsys, monkeypatch, ["--workload", "batched_x1024_tqu_nside16_lmax47", "--steps", "2", "--warmup", "3", "--no-e2e"])
    for k in REQUIRED:
        assert k in line, k
    assert line["config"]["n_batch"] == 40 and line["roofline"]["algorithmic_flop_per_unit"] == 8.0 and line["roofline"]["hbm_write_gbs"] > 0
    assert line["gpu_launches"] == 2 and "element" in line["unit"]


def test_tt_rank_of_several_takes_orbit_shards(fake_gpu, capsys, monkeypatch):
    """BASELINE configs[2] on several GPUs: orbit shards without transposed images, no exchange"""
    argv = ["--workload", "tt_nside32_lmax96", "--gpus", "8", "--steps", "2", "--warmup", "3"]
    line = _run(fake_gpu, capsys, monkeypatch, argv, rank=0, world=8)
    assert line["n_gpus"] == 8 and "symmetry orbits" in line["path"]["method"] and line["exchange"] is None
    r = line["roofline"]
    assert r["evaluated_pixel_pairs"] * 8 < 0.4 * r["stored_pixel_pairs_all_ranks"]          # 22.5 / 72 of the pairs, an eighth of them here
