"""How long does the HOST take to enqueue one encoder pass / one full generate (no device sync inside)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deephumor_b200.runtime import ops
from deephumor_b200.utils import synth
model, hp, sd = bench.build_model('lstm_labels', 'bf16')
dev = 'cuda'
imgs = torch.empty(512, 3, 224, 224, device=dev); ops.synth_images(imgs, 0, 0)
labs = synth.labels(0, 0, 512, bench.V).to(dev)
host = torch.empty(512, 3, 224, 224, pin_memory=True); host.copy_(imgs)
kw = dict(max_len=32, temperature=1.0, beam_size=5, top_k=50, noise='injected', seed=1)
with torch.no_grad():
    for _ in range(3):
        model.generate(imgs, labs, **kw)
    torch.cuda.synchronize()
    for name, inp in (('device images', imgs), ('pinned host images', host)):
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            enc = model.encoder(inp, labs)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            out = model.generate(inp, labs, **kw)
            t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
        print(f'{name}: encoder enqueue {1e3*(t1-t0):.2f} ms (done after {1e3*(t2-t0):.2f}); generate enqueue+finish() {1e3*(t3-t2):.2f} ms (done after {1e3*(t4-t2):.2f})')
