"""Data-parallel numerics on real GPUs (needs >= 2; skipped otherwise): two NCCL ranks, each with half of a batch,
must produce the same averaged gradient — and the same parameters after the fused clip + AdamW step — as one
process on the whole batch (noise keyed by the global sample index)."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    import tempo_vae_oracle as orc
    import tempo_vae_b200 as t
    from test_model_gpu import params_for
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_train.pt"), weights_only=False)
    t.seed_all(42)
    model = t.get_model(params_for(orc.TINY_CFG), dev)
    model.load_state_dict(fx["state_dict"])
    return model, orc


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import tempo_vae_b200 as t
    from tempo_vae_b200.parallel import DataParallel
    model, orc = _build(dev)
    dp = DataParallel(model, model.optimizer, bucket_mb=0.05)          # several buckets even for the tiny model
    B = 8
    x = orc.structured_batch(B, orc.TINY_CFG, seed=77)
    per = B // world
    t.seed_all(9)
    res = []
    for step in range(2):
        xl = x[rank * per:(rank + 1) * per].to(dev)
        loss, metrics = dp.get_loss(xl)
        model.optimizer.zero_grad()
        dp.backward(loss)
        grads = model.optimizer.flat_grad.clone() / world
        dp.step(max_grad_norm=1.0)
        res.append((grads.cpu(), model.optimizer.flat_param.clone().cpu(), float(loss)))
    if rank == 0:
        torch.save(res, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_process_step(tmp_path):
    import torch.multiprocessing as mp
    import tempo_vae_b200 as t
    out_path = str(tmp_path / "dp.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    dp_res = torch.load(out_path, weights_only=False)

    dev = torch.device("cuda", 0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    model, orc = _build(dev)
    x = orc.structured_batch(8, orc.TINY_CFG, seed=77).to(dev)
    t.seed_all(9)
    for step in range(2):
        loss, _ = model.get_loss(x)
        model.optimizer.zero_grad()
        loss.backward()
        g = model.optimizer.flat_grad.clone().cpu()
        model.optimizer.step(max_grad_norm=1.0)
        p = model.optimizer.flat_param.clone().cpu()
        g_dp, p_dp, loss_dp = dp_res[step]
        # per-rank loss is the mean over the local half; gradients are averaged over ranks
        rel = ((g_dp - g).norm() / g.norm()).item()
        assert rel < 2e-3, (step, rel)                      # same math, different bf16 summation order in wgrad
        assert (p_dp - p).abs().max().item() < 2.5e-4       # AdamW normalises: sign flips of ~0 gradients move <= 2 lr
        assert ((p_dp - p).norm() / p.norm()).item() < 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_cuda1_while_current_device_is_0():
    """launches must follow the device the tensors live on (the reference's torch ops work from any current
    device): build, step and encode a model on cuda:1 while the current device stays 0, same numbers as on cuda:0."""
    import tempo_vae_b200 as t
    torch.cuda.set_device(0)
    out = []
    for idx in (0, 1):
        dev = torch.device("cuda", idx)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        model, orc = _build(dev)
        x = orc.structured_batch(4, orc.TINY_CFG, seed=5).to(dev)
        t.seed_all(3)
        loss, _ = model.get_loss(x)
        model.optimizer.zero_grad()
        loss.backward()
        model.optimizer.step(max_grad_norm=1.0)
        with torch.no_grad():
            mean = model.get_latent(x).mean
        assert torch.cuda.current_device() == 0
        assert mean.device == dev and model.optimizer.flat_param.device == dev
        out.append((float(loss.detach()), model.optimizer.flat_param.cpu().clone(), mean.cpu().clone()))
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    # the C ABI refuses a pointer that lives on another device instead of faulting inside the kernel
    from tempo_vae_b200 import _lib
    a = torch.zeros(64, device="cuda:1")
    b = torch.zeros(64, dtype=torch.bfloat16, device="cuda:1")
    rc = _lib.lib.tvae_f32_to_bf16(a.data_ptr(), b.data_ptr(), 64, None, torch.cuda.current_stream(0).cuda_stream)
    assert rc != 0 and "device" in _lib.last_error()
    assert t.get_device().type == "cuda"
