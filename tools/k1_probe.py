"""ncu target: K1 alone over a device-resident text window (development aid; the launch bench.py times for `roofline`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrzip_next_b200 import Context, datagen
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mask = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = mb << 20
h = datagen.gen_text_blocks(n)
d = torch.zeros(n + 8192 + 256, dtype=torch.uint8, device="cuda")
d[256:256 + n] = torch.from_numpy(h).cuda()
tiles = (n + 511) // 512
cand = torch.empty(tiles * 512 * 2, dtype=torch.int64, device="cuda")
tc = torch.empty(tiles, dtype=torch.int32, device="cuda")
ctx = Context(0)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.k1_launch(d.data_ptr() + 256, n, mask, cand.data_ptr(), tc.data_ptr(), st)
torch.cuda.synchronize()
print("candidates", int(tc.sum().item()), "of", n)
