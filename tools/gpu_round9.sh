#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_retrieval.py -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_retrieval.log 2>&1; echo "retrieval exit $?"; tail -30 gpurun_out/pytest_retrieval.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider --deselect tests/test_retrieval.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from clip_lite_b200 import retrieval as R
from oracle import retrieval_oracle as ro
img, txt, txt2img, img2txt, ids = ro.synth_retrieval(5000, 5, 2048, seed=3, noise=30.0)
img2idx = {i: k for k, i in enumerate(ids)}
rows = [img2txt[i] for i in ids]
cols = torch.tensor([img2idx[txt2img[j]] for j in range(txt.shape[0])])
a, b = img.cuda(), txt.cuda()
for prec in ("bf16x3", "bf16"):
    for _ in range(2): R.retrieval_ranks(a, b, rows, cols, precision=prec)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): r = R.retrieval_ranks(a, b, rows, cols, precision=prec)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"COCO-5k shape (5000 x 25000, D=2048) {prec}: {dt*1e3:.2f} ms per evaluation (host CSR build included)")
t0 = time.perf_counter()
s = (a @ b.t()).cpu().numpy()
t1 = time.perf_counter()
res = ro.itm_eval(s, s.T, txt2img, img2txt, ids)
t2 = time.perf_counter()
print(f"reference route: matmul + copy to host {1e3*(t1-t0):.1f} ms, NumPy ranking (oracle restatement, not argsort) {1e3*(t2-t1):.1f} ms")
PY
