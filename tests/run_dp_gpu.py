"""Multi-GPU parity check, run as:  torchrun --nproc-per-node R tests/run_dp_gpu.py
Each rank trains on its contiguous shard of a golden batch (NCCL sync-BatchNorm, broadcast DGI negatives,
averaged gradients); rank 0 compares the gathered result with the single-process reference fixture, for the
eager path and for the CUDA-graph path (several steps)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import Golden, assert_close, grad_floor  # noqa: E402
from graph_neural_mapping_b200 import dist as gdist  # noqa: E402
from graph_neural_mapping_b200.models import GIN_InfoMaxReg  # noqa: E402


def build(g, dev, comm, graphs_on):
    c = g.cfg
    m = GIN_InfoMaxReg(c["num_layers"], c["num_mlp_layers"], c["input_dim"], c["hidden_dim"], c["output_dim"],
                       c["final_dropout"], c["learn_eps"], c["graph_pooling_type"], c["neighbor_pooling_type"], dev)
    m.load_state_dict(g.state_dict())
    m = m.to(dev)
    m.set_comm(comm)
    m.use_cuda_graphs = graphs_on
    return m


def step(model, graphs, beta, seed, dev, comm):
    model.train()
    np.random.seed(seed)
    c_logit, d_logit = model(graphs)
    labels = torch.LongTensor([x.label for x in graphs]).to(dev)
    n = len(graphs) * graphs[0].node_features.shape[1]
    d_labels = torch.cat([torch.ones(n, 1), torch.zeros(n, 1)], 0).to(dev)
    loss = torch.nn.functional.cross_entropy(c_logit, labels) + beta * \
        torch.nn.functional.binary_cross_entropy_with_logits(d_logit, d_labels)
    model.zero_grad()
    loss.backward()
    gdist.average_gradients(model, comm)
    return c_logit, d_logit, loss


def gather(t, comm):
    out = [torch.empty_like(t) for _ in range(comm.world)]
    td.all_gather(out, t.contiguous())
    return out


def check_peer_exchange(comm, dev):
    """The NVLink peer-memory all-reduce (include/gnm.h, data-parallel section) against NCCL: 200 back-to-back
    exchanges of varying length, bit-identical results on every rank, no give-ups."""
    p2p = comm.p2p
    if p2p is None:
        if os.environ.get("GNM_P2P", "1") != "0" and comm.rank == 0:
            print("peer exchange NOT available on this box (falling back to NCCL)", flush=True)
        return os.environ.get("GNM_P2P", "1") == "0"
    gen = torch.Generator(device="cpu").manual_seed(1234 + comm.rank)
    for it in range(200):
        n = [1, 2, 128, 256, 37][it % 5]
        x = (torch.randn(n, generator=gen, dtype=torch.float64) * 10.0 ** (it % 7 - 3)).to(dev)
        got = p2p.allreduce(x.clone())
        # the kernel adds the ranks' payloads in rank order: the same sequence of fp64 additions, done here by hand
        # on the NCCL-gathered inputs, must give the same bits; NCCL's own all-reduce (another order) only nearly so
        parts = gather(x, comm)
        ref = torch.zeros_like(x)
        for part in parts:
            ref = ref + part
        if not torch.equal(got, ref):
            print("rank %d: peer exchange %d differs from the rank-ordered sum" % (comm.rank, it), flush=True)
            return False
        nccl = x.clone()
        td.all_reduce(nccl)
        scale = torch.stack([q.abs() for q in parts]).sum(0)
        if not bool(((got - nccl).abs() <= 1e-14 * scale + 1e-300).all()):
            print("rank %d: peer exchange %d differs from NCCL beyond rounding" % (comm.rank, it), flush=True)
            return False
        every = gather(got, comm)
        if not all(torch.equal(every[0], e) for e in every):
            print("rank %d: peer exchange %d not bit-identical across ranks" % (comm.rank, it), flush=True)
            return False
    # the flat gradient all-reduce over peer regions (push to every peer, barrier, rank-ordered sum): odd lengths,
    # repeated calls (slot reuse), growth of the region
    for it, n in enumerate([167435, 167435, 7, 690192, 64]):
        x = torch.randn(n, generator=gen).to(dev) * (it + 1)
        parts = gather(x, comm)
        ref = torch.zeros_like(x)
        for part in parts:
            ref = ref + part                          # the kernel's order of additions
        ref = ref * (1.0 / comm.world)
        got = comm.all_reduce_mean_flat(x.clone())
        if not torch.equal(got, ref):
            print("rank %d: flat gradient all-reduce %d differs from the rank-ordered mean (max %.3e)"
                  % (comm.rank, it, float((got - ref).abs().max())), flush=True)
            return False
        every = gather(got, comm)
        if not all(torch.equal(every[0], e) for e in every):
            print("rank %d: flat gradient all-reduce %d not bit-identical across ranks" % (comm.rank, it), flush=True)
            return False
    torch.cuda.synchronize()
    if p2p.status():
        print("rank %d: a peer exchange gave up waiting" % comm.rank, flush=True)
        return False
    if comm.rank == 0:
        print("peer exchange OK: world=%d, 200 exchanges + 5 flat all-reduces, bit-identical on all ranks" % comm.world, flush=True)
    return True


def check_trainer(comm, dev, name, seed0):
    """driver.Trainer (step on libgnm's own kernels, DGI rows and gradients over peer memory) on the sharded golden batch:
    the mean of the ranks' losses must equal the single-process reference's loss, for the eager steps and the captured
    CUDA graph; afterwards every rank must hold bit-identical parameters."""
    from graph_neural_mapping_b200.driver import Trainer
    g = Golden(name)
    if g.cfg["B"] % comm.world:
        return True
    graphs = gdist.shard(g.graphs(), comm)
    model = build(g, dev, comm, True)
    model.final_dropout = 0.0
    model.train()
    tr = Trainer(model, lr=0.005, beta=g.cfg["beta"], comm=comm)
    ok = True
    for rep in range(4):
        model.load_state_dict(g.state_dict())
        np.random.seed(4242 + seed0)
        loss = tr.step(graphs).detach().reshape(1).clone()
        ls = gather(loss, comm)
        if comm.rank == 0:
            try:
                assert_close(torch.stack(ls).mean(), g.z["train/loss"], 1e-4, "Trainer loss, step %d" % rep)
            except AssertionError as e:
                print(str(e), flush=True)
                ok = False
    tr.finish()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    every = gather(flat, comm)
    if not all(torch.equal(every[0], e) for e in every):
        print("rank %d: parameters differ across ranks after Trainer steps" % comm.rank, flush=True)
        ok = False
    if comm.rank == 0 and ok:
        print("DP Trainer OK: %s world=%d" % (name, comm.world), flush=True)
    tr.release()
    model.release_graphs()
    return ok


def main():
    comm, local_rank = gdist.init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    ok = check_peer_exchange(comm, dev)
    ok = check_trainer(comm, dev, "schaefer400_b16_noeps", 2000) and ok
    ok = check_trainer(comm, dev, "mid_eps_sum_h64", 900) and ok
    # B = 16 fixtures shard over 2, 4 and 8 ranks (M = 6400 rows: the tcgen05 kernel family at world <= 2 ... the
    # small-problem kernels below 4096 rows per rank)
    for name, seed0 in [("schaefer400_b16_noeps", 2000), ("schaefer400_b16_eps", 2100), ("mid_eps_sum_h64", 900),
                        ("tiny_eps_sum", 100)]:
        g = Golden(name)
        if g.cfg["B"] % comm.world:
            continue
        graphs = gdist.shard(g.graphs(), comm)
        for graphs_on in (False, True):
            model = build(g, dev, comm, graphs_on)
            reps = 3 if graphs_on else 1          # 1st sighting eager, then capture, then replay
            for rep in range(reps):
                model.load_state_dict(g.state_dict())
                c_logit, d_logit, loss = step(model, graphs, g.cfg["beta"], 4242 + seed0, dev, comm)
            cs, ds = gather(c_logit.detach(), comm), gather(d_logit.detach(), comm)
            ls = gather(loss.detach().reshape(1), comm)
            if comm.rank == 0:
                m_local = ds[0].shape[0] // 2
                d_all = torch.cat([x[:m_local] for x in ds] + [x[m_local:] for x in ds], 0)
                assert_close(torch.cat(cs, 0), g.z["train/c_logit"], 1e-4, "c_logit")
                assert_close(d_all, g.z["train/d_logit"], 1e-4, "d_logit")
                assert_close(torch.stack(ls).mean(), g.z["train/loss"], 1e-4, "loss")
                ref = g.group("grad/")
                floor = grad_floor(ref)
                for k, p in model.named_parameters():
                    if k in ref:
                        # 3e-3 of the tensor's own max-abs from the fp64 truth (tests/test_gpu_model.py) plus the
                        # reference fixture's own distance from it (<= 1.1e-3 at N = 400)
                        assert_close(p.grad, ref[k], 4.5e-3, "grad " + k, floor=floor(k))
                if not graphs_on:
                    for k, v in g.group("buf_after/").items():
                        if "num_batches" not in k:
                            assert_close(model.state_dict()[k], v, 1e-4, k)
                print("DP parity OK: %s world=%d cuda_graphs=%s" % (name, comm.world, graphs_on), flush=True)
            # captured graphs hold references on the NCCL communicator: drop them before tearing it down
            model.release_graphs()
            del model
    torch.cuda.synchronize()
    if comm.p2p is not None and comm.p2p.status():
        print("rank %d: a peer exchange gave up waiting during the parity steps" % comm.rank, flush=True)
        ok = False
    flags = [None] * comm.world
    td.all_gather_object(flags, bool(ok))
    ok = all(flags)
    td.barrier()
    if comm.rank == 0:
        print("DP_GPU_CHECK_PASSED" if ok else "DP_GPU_CHECK_FAILED", flush=True)
    gdist.shutdown()


if __name__ == "__main__":
    main()
