"""Data-parallel host logic on CPU: world_size 2 over gloo, kernels replaced by the CPU stand-in
(tests/emul_ops.py). Each rank gets a contiguous shard of the golden batch; the union of the ranks'
outputs and the averaged gradients must equal the single-process reference at the same global batch
(sync-BatchNorm statistics, broadcast DGI negatives, reduced negative-row gradients)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, seed, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as td
    import emul_ops
    from helpers import Golden
    from graph_neural_mapping_b200 import dist as gdist, engine
    from graph_neural_mapping_b200.models import graphcnn as gmod, mlp as mlpmod, discriminator as dmod
    torch.set_num_threads(1)
    engine._ops = emul_ops
    mlpmod._ops = emul_ops
    dmod._ops = emul_ops
    engine.require_cuda = lambda dev: None
    td.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    comm = gdist.Comm(None, world, rank)
    g = Golden(name)
    c = g.cfg
    model = gmod.GIN_InfoMaxReg(c["num_layers"], c["num_mlp_layers"], c["input_dim"], c["hidden_dim"], c["output_dim"],
                                c["final_dropout"], c["learn_eps"], c["graph_pooling_type"], c["neighbor_pooling_type"],
                                torch.device("cpu"))
    model.load_state_dict(g.state_dict())
    model.set_comm(comm)
    graphs = gdist.shard(g.graphs(), comm)
    model.train()
    np.random.seed(seed)
    c_logit, d_logit = model(graphs)
    labels = torch.LongTensor([x.label for x in graphs])
    n = len(graphs) * graphs[0].node_features.shape[1]
    d_labels = torch.cat([torch.ones(n, 1), torch.zeros(n, 1)], 0)
    loss = torch.nn.functional.cross_entropy(c_logit, labels) + c["beta"] * \
        torch.nn.functional.binary_cross_entropy_with_logits(d_logit, d_labels)
    model.zero_grad()
    loss.backward()
    gdist.average_gradients(model, comm)
    out = {"c_logit": c_logit.detach().numpy(), "d_logit": d_logit.detach().numpy(), "loss": loss.detach().numpy()}
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy()
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["buf/" + k] = v.numpy()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("name,seed", [("tiny_eps_sum", 4342), ("tiny_noeps_avg", 4542), ("mid_eps_sum_h64", 5142)])
def test_two_ranks_match_single_process_reference(name, seed, tmp_path):
    sys.path.insert(0, HERE)
    from helpers import Golden, assert_close, grad_floor
    world = 2
    g = Golden(name)
    if g.cfg["B"] % world:
        pytest.skip("batch not divisible")
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, seed, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world)]
    c_logit = np.concatenate([x["c_logit"] for x in r], 0)
    m_local = r[0]["d_logit"].shape[0] // 2
    pos = np.concatenate([x["d_logit"][:m_local] for x in r], 0)
    neg = np.concatenate([x["d_logit"][m_local:] for x in r], 0)
    assert_close(c_logit, g.z["train/c_logit"], 1e-4, "c_logit")
    assert_close(np.concatenate([pos, neg], 0), g.z["train/d_logit"], 1e-4, "d_logit")
    assert_close(np.mean([x["loss"] for x in r]), g.z["train/loss"], 1e-4, "loss")
    ref = g.group("grad/")
    floor = grad_floor(ref)
    for k, v in ref.items():
        assert_close(r[0]["grad/" + k], v, 2e-3, "grad " + k, floor=floor(k))
        assert np.array_equal(r[0]["grad/" + k], r[1]["grad/" + k]), "ranks disagree on " + k
    for k, v in g.group("buf_after/").items():
        if "num_batches" in k:
            assert int(r[0]["buf/" + k]) == int(v)
        else:
            assert_close(r[0]["buf/" + k], v, 1e-4, k)
            assert_close(r[1]["buf/" + k], v, 1e-4, k)
