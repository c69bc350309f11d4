"""How long does the HOST need to enqueue one cfg2 training step (python + ctypes + launches) compared with the device time?"""
import sys
import time
import torch
sys.path.insert(0, ".")
import bench
import rbm_b200

dev = torch.device("cuda", 0)
margs = bench.model_args(str(dev), bench.CFG["dropout"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
margs.train_batch_size = B
model = rbm_b200.model_factory(margs)
trainer = rbm_b200.trainer_factory(margs, model, None, None, None, None)
model.train()
devb = [(torch.from_numpy(t).to(dev), torch.from_numpy(l).to(dev)) for t, l in bench.make_batches(4, B, seed=100)]
for i in range(5):
    trainer.train_step(devb[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(30):
    trainer.train_step(devb[i % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("B=%d " % B + "host enqueue %.2f ms/step, wall incl. device drain %.2f ms/step" % ((t1 - t0) * 1000 / 30, (t2 - t0) * 1000 / 30))
