"""Peer-memory store bandwidth over NVLink as the pushing kernels see it (diagnostic; torchrun, >= 2 GPUs):
SM stores (an elementwise kernel writing into the peer's symmetric buffer) in contiguous form and as 64 / 128 / 256-byte
row segments of a pitched panel, every rank writing to its right neighbour at once; plus the copy engine for reference.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_store_bw.py
"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    gname = dist.group.WORLD.group_name
    try:
        symm.enable_symm_mem_for_group(gname)
    except Exception:       # noqa: BLE001
        pass
    rows, pitch = 524288, 224
    n = rows * pitch
    buf = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, gname)
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    src = torch.randn(n, device=dev)
    res = {}

    def timed(tag, fn, nbytes, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[tag] = round(nbytes / ms / 1e6, 1)

    timed("copy_engine_contiguous", lambda: peer.copy_(src), n * 4)
    timed("sm_store_contiguous", lambda: torch.add(src, 1.0, out=peer), n * 4)
    P, S = peer.view(rows, pitch), src.view(rows, pitch)
    for wcols in (16, 32, 64, 112):
        nseg = pitch // wcols
        def f(wcols=wcols, nseg=nseg):
            for k in range(0, nseg, 2):                    # every other segment: the gaps stay unwritten
                torch.add(S[:, k * wcols:(k + 1) * wcols], 1.0, out=P[:, k * wcols:(k + 1) * wcols])
        timed(f"sm_store_{wcols * 4}B_segments", f, rows * wcols * 4 * len(range(0, nseg, 2)))
    timed("local_sm_store_contiguous", lambda: torch.add(src, 1.0, out=buf), n * 4)
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        for r, o in enumerate(out):
            print(f"rank {r} -> {(r + 1) % world} GB/s:", o)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
