"""cProfile of the HOST side of the train step (Python + ctypes enqueue cost), B small so the GPU never back-pressures."""
import cProfile, os, pstats, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import tempo_vae_b200 as t

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
t.seed_all(42)
model = t.get_model(bench.DEFAULT_MODEL, dev)
trainer = t.Trainer(model, model.optimizer, dev, tempfile.mkdtemp(prefix="tvae_hp_"))
x = bench.synthetic_batch(torch, B, (1028, 64, 64), dev, seed=0)
for _ in range(3):
    trainer.train_step_device(x)
    trainer.step = 1
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10):
    trainer.train_step_device(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue: {(t1 - t0) * 100:.2f} ms/step at B={B}")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    trainer.train_step_device(x)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
