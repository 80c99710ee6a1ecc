import ctypes, sys, os
sys.path.insert(0, os.getcwd())
import torch
from drsa_audio_b200 import _lib as L
torch.zeros(1).cuda()
for d in (128, 256):
    for sp in (0, 1):
        out = (ctypes.c_int * 5)()
        print(d, sp, L.lib().drsa_debug_tc_kernel_attrs(d, sp, out), list(out))
