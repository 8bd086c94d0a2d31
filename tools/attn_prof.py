import ctypes as C, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from aurora_b200 import _native as N
from aurora_b200.engine import to_bf16_bits
lib = N.load()
rng = np.random.default_rng(7)
n_seq, heads = 64, 12
hidden = heads * 64
lens = np.clip(np.rint(rng.normal(384, 96, n_seq)), 16, 512).astype(np.int64)
if os.environ.get('AUR_LEN'):
    lens[:] = int(os.environ['AUR_LEN'])
if os.environ.get('AUR_NSEQ'):
    n_seq = int(os.environ['AUR_NSEQ']); lens = np.resize(lens, n_seq)
cu = np.zeros(n_seq + 1, np.int32); cu[1:] = np.cumsum(lens)
T = int(cu[-1])
qkv = to_bf16_bits((rng.standard_normal((T, 3 * hidden)) * 1.0).astype(np.float32))
out = np.zeros((T, hidden), dtype=np.uint16)
ms = C.c_float()
N.check(lib.aur_debug_attention(0, qkv.ctypes.data_as(C.c_void_p), cu.ctypes.data_as(C.c_void_p), n_seq, heads, hidden, out.ctypes.data_as(C.c_void_p), C.byref(ms)))
print("attention", T, "tokens", ms.value * 1e3, "us")
