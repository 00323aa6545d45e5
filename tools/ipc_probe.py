"""Feasibility probe: CUDA IPC memory handles + P2P stores between two ranks of one box.
torchrun --nproc-per-node 2 tools/ipc_probe.py"""
import ctypes as C, os, sys
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
try:
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
except OSError:
    pass
class Handle(C.Structure):
    _fields_ = [("r", C.c_ubyte * 64)]


rt.cudaIpcGetMemHandle.argtypes = [C.POINTER(Handle), C.c_void_p]
rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), Handle, C.c_uint]     # the handle is passed BY VALUE
ptr = C.c_void_p()
assert rt.cudaMalloc(C.byref(ptr), 4096) == 0
rt.cudaMemset(ptr, 0, 4096)
h = Handle()
rc = rt.cudaIpcGetMemHandle(C.byref(h), ptr)
print(f"[rank {rank}] cudaIpcGetMemHandle rc={rc}", flush=True)
mine = torch.tensor(list(bytes(h.r)), dtype=torch.uint8, device="cuda")
allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
dist.all_gather_into_tensor(allh, mine)
allh = bytes(allh.cpu().numpy().tobytes())
peer = (rank + 1) % world
ph = Handle.from_buffer_copy(allh[peer * 64:(peer + 1) * 64])
pp = C.c_void_p()
rc = rt.cudaIpcOpenMemHandle(C.byref(pp), ph, 1)      # cudaIpcMemLazyEnablePeerAccess
print(f"[rank {rank}] cudaIpcOpenMemHandle(peer {peer}) rc={rc} ptr={pp.value}", flush=True)
if rc == 0:
    val = (C.c_uint32 * 4)(100 + rank, 1, 2, 3)
    rc = rt.cudaMemcpy(pp, val, 16, 1)
    print(f"[rank {rank}] write to peer rc={rc}", flush=True)
dist.barrier()
torch.cuda.synchronize()
out = (C.c_uint32 * 4)()
rt.cudaMemcpy(out, ptr, 16, 2)
print(f"[rank {rank}] my buffer now holds {list(out)} (expect {100 + (rank - 1) % world} from the peer)", flush=True)
dist.destroy_process_group()
