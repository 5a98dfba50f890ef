"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes drive the GradBucketer exactly as
the engine's backward does (parameters report in reverse order, a never-used parameter never reports)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeParam:
    pass


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tempo_vae_b200.parallel import GradBucketer
    sizes = [1, 300, 64, 5000, 17, 2048, 33]          # element counts; param 4 never gets a gradient
    params = [FakeParam() for _ in sizes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 15) // 16 * 16
    flat = torch.zeros(total)
    ranges = [(p, o, o + n) for p, o, n in zip(params, offs, sizes)]
    bk = GradBucketer(flat, ranges, bucket_bytes=4 * 1000)
    assert len(bk.buckets) >= 3 and bk.buckets[-1][1] == total
    results = []
    for it in range(3):                                # iteration 0 learns which parameters report
        flat.zero_()
        for i in reversed(range(len(params))):         # backward order
            if i == 4:
                continue
            o, n = offs[i], sizes[i]
            flat[o:o + n] = (rank + 1) * (i + 1) + it
            bk.ready(params[i])
        if it > 0:
            assert any(bk.launched), "buckets must be launched from the ready hook after the first iteration"
        bk.finish()
        results.append(flat.clone())
    ok = True
    for it, r in enumerate(results):
        for i, (o, n) in enumerate(zip(offs, sizes)):
            exp = 0.0 if i == 4 else sum((rk + 1) * (i + 1) + it for rk in range(world))
            ok = ok and bool((r[o:o + n] == exp).all())
            pad = r[o + n:(o + n + 15) // 16 * 16]
            ok = ok and bool((pad == 0).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_grad_bucketer_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(out) == [(0, True), (1, True)]
