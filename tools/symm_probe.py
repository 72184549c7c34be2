import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:2], "sig size", hdl.signal_pad_size)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.uint8)
    print(rank, "peer view", peer[:4].tolist())
    hdl.barrier()
except Exception as ex:
    import traceback; traceback.print_exc()
dist.destroy_process_group()
