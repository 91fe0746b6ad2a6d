import ctypes, torch, sys
sys.path.insert(0, '/root/repo')
from attentionshift_b200 import lib
torch.cuda.init(); torch.zeros(1).cuda()
L = lib.load()
r, s, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
print('occ', L.as_mean_shift_v2_occupancy(ctypes.byref(r), ctypes.byref(s), ctypes.byref(d)), 'regs', r.value, 'static', s.value, 'dyn', d.value)
p = torch.cuda.get_device_properties(0)
print(p.name, 'smem/SM', p.shared_memory_per_multiprocessor, 'smem/block optin', p.shared_memory_per_block_optin, 'regs/SM', p.regs_per_multiprocessor)
