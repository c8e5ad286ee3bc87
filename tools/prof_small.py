"""ncu target: one rzip of a small text window (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrzip_next_b200 import Context, datagen
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = datagen.gen_text(mb << 20)
ctx = Context(0)
s0, s1, st, vr = ctx.rzip_chunk(d, 7)
print(len(s0), len(s1), st["lookups"], st["ms_rzip"])
