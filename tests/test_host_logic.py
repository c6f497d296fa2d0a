"""CPU tests of the host-side logic: reference-compatible Python surface, padded shapes, error
behaviour without a GPU (fails loudly -- no CPU fallback), row sharding and the setup-time weight
broadcast over gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_surface_matches_reference_names():
    import qutlass
    import qutlass.utils
    for name in ("matmul_mxf4_bf16_tn", "matmul_nvf4_bf16_tn", "fusedQuantizeMx", "fusedQuantizeNv",
                 "matmul_ada_mxf4_bf16_tn", "matmul_mxf8_bf16_tn", "backward_t_bf16", "mxfp4_transpose_mxfp8"):
        assert callable(getattr(qutlass, name))
    for name in ("to_blocked", "get_padded_shape_mx", "get_padded_shape_nv", "pad_to_block", "ceil_div"):
        assert callable(getattr(qutlass.utils, name))
    import inspect
    sig = inspect.signature(qutlass.fusedQuantizeMx)
    assert sig.parameters["method"].default == "quest" and sig.parameters["return_mask"].default is False
    assert inspect.signature(qutlass.fusedQuantizeNv).parameters["method"].default == "abs_max"
    assert inspect.signature(qutlass.matmul_mxf4_bf16_tn).parameters["backend"].default == "cutlass"
    assert inspect.signature(qutlass.utils.to_blocked).parameters["use_triton_kernel"].default is False
    # every schema of the reference's op library (qutlass/csrc/bindings.cpp:498-515)
    for op in ("matmul_mxf4_bf16_tn", "matmul_nvf4_bf16_tn", "matmul_ada_mxf4_bf16_tn", "matmul_mxf8_bf16_tn",
               "matmul_mxf8_bf16_nn", "fusedQuantizeMxQuest", "fusedQuantizeMxAbsMax", "fusedQuantizeNvQuest",
               "fusedQuantizeNvAbsMax", "fusedQuantizeMxQuestWithMask", "backward_t_bf16", "backward_qt_bf16",
               "backward_bf16_square_double_mxfp8", "mxfp4_transpose_mxfp8"):
        assert hasattr(torch.ops._qutlass_C, op)


def test_padded_shapes_match_reference(golden):
    from qutlass_b200.utils import get_padded_shape_mx, get_padded_shape_nv
    assert get_padded_shape_mx(torch.empty(3, 200, 4096)) == tuple(golden["padded_mx_3x200x4096"])
    assert get_padded_shape_nv(torch.empty(3, 200, 4096)) == tuple(golden["padded_nv_3x200x4096"])
    assert get_padded_shape_mx(torch.empty(1, 96)) == tuple(golden["padded_mx_1x96"])
    assert get_padded_shape_nv(torch.empty(1, 96)) == tuple(golden["padded_nv_1x96"])


def test_no_cpu_fallback():
    """the product path must fail loudly without a CUDA device, never compute on the CPU."""
    import qutlass_b200 as Q
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        Q.fusedQuantizeMx(x, torch.eye(32, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="CUDA"):
        Q.to_blocked(torch.zeros(128, 4, dtype=torch.uint8))
    a = torch.zeros(4, 32, dtype=torch.uint8)
    sf = torch.zeros(512, dtype=torch.float8_e8m0fnu)
    with pytest.raises(RuntimeError, match="CUDA"):
        Q.matmul_mxf4_bf16_tn(a, a, sf, sf, torch.ones(1))


def test_backend_and_method_validation_happen_before_any_device_work():
    import qutlass_b200 as Q
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        Q.fusedQuantizeMx(x, torch.eye(32, dtype=torch.bfloat16), method="bogus")
    with pytest.raises(ValueError):
        Q.fusedQuantizeNv(x, torch.eye(32, dtype=torch.bfloat16), torch.ones(1), method="bogus")
    with pytest.raises(ValueError):
        Q.matmul_nvf4_bf16_tn(x, x, x, x, x, backend="bogus")
    with pytest.raises(ImportError):
        Q.matmul_nvf4_bf16_tn(x, x, x, x, x, backend="flashinfer")


@pytest.mark.parametrize("m,world", [(16384, 8), (16384, 4), (4096, 2), (1000, 3), (1, 2), (128, 8)])
def test_shard_rows_partition(m, world):
    from qutlass_b200.sharding import shard_rows
    covered = 0
    for r in range(world):
        start, rows = shard_rows(m, world, r)
        assert start == min(covered, m) or rows == 0
        assert start % 128 == 0 or rows == 0
        covered += rows
    assert covered == m
    if m % (128 * world) == 0:
        assert all(shard_rows(m, world, r)[1] == m // world for r in range(world))


def _gloo_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import helpers as H
    from qutlass_b200.sharding import broadcast_weights, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, n, k = 300, 64, 256
    # rank 0 owns the quantised weights; the others start with garbage and receive the ONE broadcast
    bq, bsf = H.random_fp4_operand(n, k, "mx", seed=2, sf_mode="narrow")
    wq = torch.from_numpy(bq.copy()) if rank == 0 else torch.zeros(n, k // 2, dtype=torch.uint8)
    wsf = torch.from_numpy(H.blocked_sf(bsf)).view(torch.float8_e8m0fnu) if rank == 0 else \
        torch.zeros(128 * 8, dtype=torch.float8_e8m0fnu)
    broadcast_weights(wq, wsf, src=0)
    assert np.array_equal(wq.numpy(), bq)
    assert np.array_equal(wsf.view(torch.uint8).numpy(), H.blocked_sf(bsf))
    # each rank multiplies only its own activation rows (oracle stands in for the GPU kernel on CPU)
    aq, asf = H.random_fp4_operand(m, k, "mx", seed=1, sf_mode="narrow")
    start, rows = shard_rows(m, world, rank)
    sf_back = O.from_blocked(wsf.view(torch.uint8).numpy(), 128, 8)[:n, : k // 32]
    part = H.gemm_oracle_bits(aq[start:start + rows], asf[start:start + rows], wq.numpy(), sf_back, "mx")
    np.save(os.path.join(tmp, f"part{rank}.npy"), part)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_row_sharding_reassembles_full_product(tmp_path):
    import socket
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"part{r}.npy") for r in range(world)]
    aq, asf = H.random_fp4_operand(300, 256, "mx", seed=1, sf_mode="narrow")
    bq, bsf = H.random_fp4_operand(64, 256, "mx", seed=2, sf_mode="narrow")
    full = H.gemm_oracle_bits(aq, asf, bq, bsf, "mx")
    np.testing.assert_array_equal(np.concatenate(parts, axis=0), full)


def test_bench_reference_arm_prints_contract_line():
    """--impl reference must print one JSON line with the contract keys (tiny run on the CPU)."""
    import json
    import subprocess
    env = dict(os.environ, B200Q_BENCH_TINY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0
