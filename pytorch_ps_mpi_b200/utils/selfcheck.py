"""Multi-GPU self-check used by ``__graft_entry__.smoke()`` when more than one GPU is visible."""
from __future__ import annotations

import torch


def multi_gpu_smoke(rank: int, size: int) -> None:
    """One rank per GPU: two sync-PS steps of a small bf16 MLP through the device engine — the gather runs over peer
    memory (``multimem.ld_reduce`` at N >= 4, rank-ordered P2P below) and the broadcast is ONE ``multimem.st`` per 16 bytes
    through the switch when the fabric has NVLS.  All ranks must end bit-identical."""
    import pytorch_ps_mpi_b200 as ps
    from pytorch_ps_mpi_b200.models import mnist_mlp
    w = ps.runtime.init()
    assert (w.rank, w.size) == (rank, size)
    dev = w.device
    torch.manual_seed(0)
    model = mnist_mlp(hidden=256).to(dev).bfloat16()
    opt = ps.SGD(model.named_parameters(), model.parameters(), lr=0.05, momentum=0.9, mode="ps", engine="device")
    eng = opt._engine
    for s in range(2):
        g = torch.Generator().manual_seed(10 * s + rank)
        x = torch.randn(32, 784, generator=g).to(dev).bfloat16()
        y = torch.randint(0, 10, (32,), generator=g).to(dev)
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(model(x).float(), y).backward()
        opt.step()
    eng.check()
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).cpu()
    allp = w.all_gather_object(flat)
    assert all(torch.equal(f, allp[0]) for f in allp), "ranks diverged"
    if rank == 0:
        print(f"multi-GPU smoke ok: ranks={size} multicast={eng.arena.has_multicast} "
              f"bcast={ {0: 'local', 1: 'unicast-p2p', 2: 'multimem.st'}[eng.bcast] } "
              f"reduce={ {0: 'p2p', 1: 'multimem.ld_reduce'}[eng.reduce] } chunks={eng.nchunks}", flush=True)
    opt.close()
