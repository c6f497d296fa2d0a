"""Multi-GPU parity on real hardware (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a one-GPU box).  One process per GPU, NCCL: weights quantised on rank 0 and broadcast once
(qutlass_b200.sharding.broadcast_weights), every rank quantises + multiplies its own row shard (shard_rows), an UN-TIMED
all_gather brings the shards to rank 0, which checks them against its own full-M product bit for bit."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA required", allow_module_level=True)


def _worker(rank, world, port, fmt, m, n, k):
    sys.path.insert(0, ROOT)
    # One quantiser kernel for every shard size: by default inputs below 8 M elements take the butterfly kernel and larger ones
    # the tcgen05 kernel, which differ in fp32 summation order of the rotation (<= 1e-5 of the codes) -- a 512-row shard and
    # the 4096-row full tensor would then not be bit-comparable (observed at 8 GPUs: 0.12 % of the outputs one ulp apart).
    os.environ["B200Q_QUANT_TC"] = "1"
    import torch.distributed as dist
    import qutlass_b200 as Q
    from qutlass_b200.sharding import broadcast_weights, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    idx = torch.arange(128)
    bits = idx[:, None] & idx[None, :]
    par = torch.zeros_like(bits)
    while bits.any():
        par ^= bits & 1
        bits = bits >> 1
    R = ((1.0 - 2.0 * par.double()) * 128 ** -0.5).to(torch.bfloat16).to(dev)
    gs = torch.tensor([1.0], device=dev)
    alpha = torch.tensor([1.0 / 3.0], device=dev)
    group = 32 if fmt == "mx" else 16
    sf_dt = torch.float8_e8m0fnu if fmt == "mx" else torch.float8_e4m3fn
    fq = (lambda t: Q.fusedQuantizeMx(t, R, method="abs_max")) if fmt == "mx" else (lambda t: Q.fusedQuantizeNv(t, R, gs, method="abs_max"))
    mm = Q.matmul_mxf4_bf16_tn if fmt == "mx" else Q.matmul_nvf4_bf16_tn
    wq = torch.empty(n, k // 2, dtype=torch.uint8, device=dev)
    wsf = torch.empty((n + 127) // 128 * 128 * ((k // group + 3) // 4 * 4), dtype=sf_dt, device=dev)
    if rank == 0:
        w = torch.randn(n, k, dtype=torch.bfloat16, device=dev, generator=torch.Generator(dev).manual_seed(1)) * 25
        q_, s_ = fq(w)
        wq.copy_(q_)
        wsf.copy_(Q.to_blocked(s_))
    else:
        wq.fill_(0x55)
    broadcast_weights(wq, wsf, src=0)
    # every rank derives the SAME activations from the seed, then touches only its own rows
    x = torch.randn(m, k, dtype=torch.bfloat16, device=dev, generator=torch.Generator(dev).manual_seed(2)) * 25
    s0, rows = shard_rows(m, world, rank)
    xq, xsf = fq(x[s0:s0 + rows].contiguous())
    part = mm(xq, wq, Q.to_blocked(xsf), wsf, alpha, static_weights=True)
    sizes = [shard_rows(m, world, r)[1] for r in range(world)]
    gathered = [torch.empty(sz, n, dtype=torch.bfloat16, device=dev) for sz in sizes]
    if len(set(sizes)) == 1:
        dist.all_gather(gathered, part)
    else:                                   # ragged shards: pad to the largest
        mx = max(sizes)
        pad = torch.zeros(mx, n, dtype=torch.bfloat16, device=dev)
        pad[:rows] = part
        bufs = [torch.empty(mx, n, dtype=torch.bfloat16, device=dev) for _ in range(world)]
        dist.all_gather(bufs, pad)
        gathered = [b[:sz] for b, sz in zip(bufs, sizes)]
    if rank == 0:
        xq_f, xsf_f = fq(x)
        full = mm(xq_f, wq, Q.to_blocked(xsf_f), wsf, alpha)
        torch.cuda.synchronize()
        got = torch.cat(gathered, dim=0)
        assert got.shape == full.shape
        assert torch.equal(got, full), (fmt, m, n, k, (got != full).float().mean().item())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fmt,m,n,k", [("mx", 16384, 28672, 8192), ("nv", 4096, 14336, 4096), ("mx", 1000, 1024, 512)])
def test_row_shards_on_n_gpus_equal_the_full_product(fmt, m, n, k):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, fmt, m, n, k), nprocs=world, join=True)
