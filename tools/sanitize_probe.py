"""Small whole-pipeline run for compute-sanitizer (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrzip_next_b200 import BACKEND_LZMA, Context, datagen, make_params
kb = int(sys.argv[1]) if len(sys.argv) > 1 else 96
d = datagen.gen_text(kb << 10)
ctx = Context(0)
arc = ctx.compress(d, make_params(backend=BACKEND_LZMA, threads=8))
print("archive", len(arc))
