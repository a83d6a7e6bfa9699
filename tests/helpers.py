"""Shared test helpers: golden-fixture loading and tolerance checks."""
import glob
import json
import os

import numpy as np
import torch

from graph_neural_mapping_b200.synth import SynthGraph

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden(object):
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.cfg = json.loads(str(self.z["config"]))
        self.node_counts = self.z["node_counts"].tolist()
        self.labels = self.z["labels"].tolist()
        self.perm = self.z["perm"]

    def graphs(self):
        out = []
        eo = self.z["edge_off"]
        ec = self.z["edge_cat"]
        need_nb = self.cfg["neighbor_pooling_type"] == "max"
        for i, n in enumerate(self.node_counts):
            em = torch.from_numpy(ec[:, eo[i]:eo[i + 1]].copy())
            out.append(SynthGraph(n, self.labels[i], em, torch.eye(n, dtype=torch.float32), with_neighbors=need_nb))
        return out

    def state_dict(self):
        return {k[len("state/"):]: torch.from_numpy(self.z[k].copy()) for k in self.z.files if k.startswith("state/")}

    def group(self, prefix):
        return {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}

    def state_after_train(self):
        sd = self.state_dict()
        for k, v in self.group("buf_after/").items():
            sd[k] = torch.from_numpy(v.copy())
        return sd


def rel_err(a, b):
    """max|a-b| / max(|b|max, tiny): error relative to the tensor's scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-30)
    return float(np.max(np.abs(a - b))) / scale if b.size else 0.0


def assert_close(a, b, tol, what="", floor=0.0):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert (nan_a == nan_b).all(), "%s: NaN pattern differs" % what
    a = np.where(nan_a, 0.0, a)
    b = np.where(nan_b, 0.0, b)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, floor, 1e-30)
    err = float(np.max(np.abs(a - b))) / scale if b.size else 0.0
    assert err <= tol, "%s: scaled error %.3e > %.1e" % (what, err, tol)
    return err


def grad_floor(grads, frac=1e-2):
    """Gradients that are mathematically zero (a Linear bias feeding a train-mode BatchNorm)
    are pure rounding noise in the reference; compare every gradient on a scale no smaller
    than `frac` x the largest gradient entry of the model."""
    m = 0.0
    for v in grads.values():
        if v is not None and np.size(v):
            m = max(m, float(np.max(np.abs(np.asarray(v, dtype=np.float64)))))
    return frac * m
