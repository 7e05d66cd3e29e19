"""Stage-0 probe: P2P capability, torch symmetric memory + NVLS multicast
availability, peer copy bandwidth."""
import os, time, torch, torch.distributed as dist
r = int(os.environ["RANK"]); w = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
print(r, "can_access_peer", [torch.cuda.can_device_access_peer(r, p) for p in range(w) if p != r], flush=True)
try:
    import torch.distributed._symmetric_memory as sm
    t = sm.empty(1 << 20, dtype=torch.float32, device=f"cuda:{r}")
    h = sm.rendezvous(t, dist.group.WORLD.group_name)
    print(r, "symm_mem ok: multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None,
          "signal_pad", len(h.signal_pad_ptrs), flush=True)
except Exception as e:
    print(r, "symm_mem failed:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
