"""Intra-node exchange bandwidth on this box: NCCL all-gather, NCCL send/recv (the halo exchange's primitive) and direct
peer copies through torch symmetric memory (copy engines, no SMs).
    [NCCL_...=..] torchrun --nproc-per-node N tools/dist_bw_check.py [MB=1024]"""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nel = mb * (1 << 20) // 4
    x = torch.randn(nel, device=dev)
    full = torch.empty(world * nel, device=dev)

    def timed(fn, steps=5, warm=2):
        for _ in range(warm):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    ag = timed(lambda: dist.all_gather_into_tensor(full, x))

    def p2p():
        ops = []
        for j in range(world):
            if j != rank:
                ops.append(dist.P2POp(dist.irecv, full[j * nel:(j + 1) * nel], j))
                ops.append(dist.P2POp(dist.isend, x, j))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    pp = timed(p2p)
    gb_in = (world - 1) * nel * 4 / 1e9
    msg = "world=%d msg=%d MB env={%s} | all-gather %.2f ms (%.0f GB/s in per rank) | send/recv all-to-all %.2f ms (%.0f GB/s)" % (
        world, mb, ",".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("NCCL_")), ag, gb_in / ag * 1e3, pp, gb_in / pp * 1e3)
    try:
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(world * nel, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        peers = [hdl.get_buffer(j, (world * nel,), torch.float32) for j in range(world)]

        def push():
            hdl.barrier()
            for j in range(world):
                if j != rank:
                    peers[j][rank * nel:(rank + 1) * nel].copy_(x, non_blocking=True)
            hdl.barrier()
        sm = timed(push)
        ok = True
        push()
        torch.cuda.synchronize()
        dist.all_gather_into_tensor(full, x)
        for j in range(world):
            if j != rank:
                ok &= bool(torch.equal(buf[j * nel:(j + 1) * nel], full[j * nel:(j + 1) * nel]))
        msg += " | symmetric-memory peer copies %.2f ms (%.0f GB/s out per rank, data ok=%s)" % (sm, gb_in / sm * 1e3, ok)
    except Exception as ex:
        msg += " | symmetric memory unavailable: %s: %s" % (type(ex).__name__, str(ex)[:200])
    if rank == 0:
        print(msg, flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
