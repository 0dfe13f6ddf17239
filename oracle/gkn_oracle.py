"""CPU oracle for the GKN hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker or the timed CPU baseline.  Nothing under
``molecular_dynamics_neural_operator_b200/`` imports it; the product path has no CPU fallback.

It restates, in plain PyTorch/NumPy/SciPy on the CPU, what the reference computes on this path.
Each function cites the reference lines it follows (paths relative to /root/reference).

Pinning status: the reference ships no tests, golden vectors or fixtures for this path
 and cannot be imported as-is here (torch_geometric, h5py, matplotlib, imageio and
mdlearn are absent).  The oracle is instead pinned against the reference's OWN source executed in
this container with those third-party modules stubbed (``tests/golden/make_golden.py``: the real
NNConv_old.message/update, DenseNet, KernelNN, LpLoss and construct_pairdata code runs; only
torch_geometric's MessagePassing.propagate / mean aggregation is a stated restatement).  The
resulting vectors live in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this
module against them.  The PyG aggregation semantics themselves remain "parity unpinned" by any
upstream artefact (torch_geometric is not vendored and its version is not pinned by the reference).
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# Edge MLP: DenseNet([ker_in, K, K, w*w], ReLU)          graph_kernel.py:217-242, built at :271
# --------------------------------------------------------------------------------------------
def dense_net(edge_attr: torch.Tensor, net: dict) -> torch.Tensor:
    """layers.0 Linear -> ReLU -> layers.2 Linear -> ReLU -> layers.4 Linear (no out nonlinearity,
    normalize=False).  ``net`` holds 'layers.{0,2,4}.{weight,bias}'.  graph_kernel.py:239-242."""
    h = F.relu(F.linear(edge_attr, net["layers.0.weight"], net["layers.0.bias"]))
    h = F.relu(F.linear(h, net["layers.2.weight"], net["layers.2.bias"]))
    return F.linear(h, net["layers.4.weight"], net["layers.4.bias"])


def edge_hidden(edge_attr: torch.Tensor, net: dict) -> torch.Tensor:
    """h_e = output of layers.0-.3 (the part shared by every application).  graph_kernel.py:239-242."""
    h = F.relu(F.linear(edge_attr, net["layers.0.weight"], net["layers.0.bias"]))
    return F.relu(F.linear(h, net["layers.2.weight"], net["layers.2.bias"]))


# --------------------------------------------------------------------------------------------
# NNConv_old                                               graph_kernel.py:125-214
# --------------------------------------------------------------------------------------------
def scatter_mean(msg: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG aggr='mean' (graph_kernel.py:272-273, :169): sum at `index`, divided by
    clamp(count, min=1); nodes without incoming edges get 0.  [restated PyG semantics]"""
    out = torch.zeros((num_nodes, msg.size(1)), dtype=msg.dtype)
    out.index_add_(0, index, msg)
    cnt = torch.zeros(num_nodes, dtype=msg.dtype)
    cnt.index_add_(0, index, torch.ones(index.numel(), dtype=msg.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


def nnconv_forward(x, edge_index, edge_attr, net: dict, root, bias) -> torch.Tensor:
    """NNConv_old.forward -> propagate -> message -> mean -> update.  graph_kernel.py:194-209.
    flow = source_to_target: x_j = x[edge_index[0]], aggregated at edge_index[1]."""
    w_in, w_out = root.shape
    x_j = x.index_select(0, edge_index[0])  # PyG gather
    weight = dense_net(edge_attr, net).view(-1, w_in, w_out)  # :201
    msg = torch.matmul(x_j.unsqueeze(1), weight).squeeze(1)  # :202
    aggr_out = scatter_mean(msg, edge_index[1], x.size(0))
    aggr_out = aggr_out + torch.mm(x, root)  # :206
    aggr_out = aggr_out + bias  # :208
    return aggr_out


def conv_stack(x, edge_index, edge_attr, net: dict, root1, bias1, root2, bias2, depth: int):
    """depth x relu(conv1) then depth x relu(conv2), conv1.net is conv2.net.
    graph_kernel.py:271-273, :299-302."""
    for _ in range(depth):
        x = F.relu(nnconv_forward(x, edge_index, edge_attr, net, root1, bias1))
    for _ in range(depth):
        x = F.relu(nnconv_forward(x, edge_index, edge_attr, net, root2, bias2))
    return x


def nnconv_forward_chunked(x, edge_index, edge_attr, net: dict, root, bias, chunk: int = 4096):
    """Same result as nnconv_forward, evaluating the per-edge kernel in edge chunks so that large
    graphs fit in host memory (the [E,w,w] tensor is 4.8 GB at config 3)."""
    w_in, w_out = root.shape
    E = edge_index.size(1)
    msg = torch.empty((E, w_out), dtype=x.dtype)
    for e0 in range(0, E, chunk):
        sl = slice(e0, min(E, e0 + chunk))
        weight = dense_net(edge_attr[sl], net).view(-1, w_in, w_out)
        msg[sl] = torch.matmul(x.index_select(0, edge_index[0, sl]).unsqueeze(1), weight).squeeze(1)
    out = scatter_mean(msg, edge_index[1], x.size(0))
    return out + torch.mm(x, root) + bias


# --------------------------------------------------------------------------------------------
# KernelNN                                                 graph_kernel.py:245-309
# --------------------------------------------------------------------------------------------
def _sub(sd: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def lstm_window(x_position, sd: dict, window_size: int, num_residues: int) -> torch.Tensor:
    """The LSTM prologue exactly as written (graph_kernel.py:278-288): x_position [B*W, R, 3] is
    reshaped to [B, W, R, 3], swapped to [W, B, R, 3]; each window step feeds [B, R, 3] to
    nn.LSTM(3, 3) as (seq_len=B, batch=R) with the hidden state carried across window steps."""
    dt = x_position.dtype
    w_ih, w_hh = sd["lstm.weight_ih_l0"], sd["lstm.weight_hh_l0"]
    b_ih, b_hh = sd["lstm.bias_ih_l0"], sd["lstm.bias_hh_l0"]
    x = x_position.reshape(-1, window_size, num_residues, 3)
    x = torch.swapaxes(x, 0, 1)
    h = torch.zeros(num_residues, 3, dtype=dt)
    c = torch.zeros(num_residues, 3, dtype=dt)
    out = None
    for i in x:  # i: [B, R, 3] consumed by nn.LSTM as (seq_len=B, batch=R, input=3)
        steps = []
        for t in range(i.size(0)):  # nn.LSTM cell, gate order (i, f, g, o)
            gates = F.linear(i[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
            gi, gf, gg, go = gates.chunk(4, dim=1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
            steps.append(h)
        out = torch.stack(steps)  # [B, R, 3]; (h, c) carry over to the next window step
    return F.linear(out, sd["lstm_fc.weight"], sd["lstm_fc.bias"])  # :288


def kernelnn_forward(sd: dict, data, depth: int, window_size: int, num_residues: int,
                     return_latent: bool = False):
    """KernelNN.forward (graph_kernel.py:277-309) from a reference-layout state dict."""
    x = lstm_window(data.x_position, sd, window_size, num_residues)
    emb = F.embedding(data.x_aminoacid, sd["emb.weight"])  # :292
    x = x.reshape(emb.shape[0], -1)  # :293
    x = torch.cat((emb, x), dim=1)  # :296
    x = F.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))  # :298
    net = _sub(sd, "conv1.net.")
    x = conv_stack(x, data.edge_index, data.edge_attr, net, sd["conv1.root"], sd["conv1.bias"],
                   sd["conv2.root"], sd["conv2.bias"], depth)  # :299-302
    latent = x.clone()
    out = F.linear(x, sd["fc2.weight"], sd["fc2.bias"])  # :305
    return [out, latent] if return_latent else out


# --------------------------------------------------------------------------------------------
# LpLoss.rel                                               graph_kernel.py:105-119
# --------------------------------------------------------------------------------------------
def lp_loss_rel(x: torch.Tensor, y: torch.Tensor, size_average: bool = False) -> torch.Tensor:
    """sum_b ||x_b - y_b||_2 / ||y_b||_2 over x.view(B, -1); size_average=False at :547."""
    n = x.size(0)
    diff = torch.norm(x.reshape(n, -1) - y.reshape(n, -1), 2, 1)
    ynorm = torch.norm(y.reshape(n, -1), 2, 1)
    return torch.mean(diff / ynorm) if size_average else torch.sum(diff / ynorm)


# --------------------------------------------------------------------------------------------
# construct_pairdata                                       graph_kernel.py:362-393
# --------------------------------------------------------------------------------------------
def construct_pairdata_arrays(x_position: np.ndarray, threshold: float = 8.0):
    """Radius graph of the LAST window frame, via the very SciPy calls the reference makes
    (graph_kernel.py:363-368), plus the per-edge attribute loop (:372-379) vectorised.
    Returns (edge_index int64 [2,E], edge_attr float32 [E,6])."""
    from scipy.sparse import coo_matrix
    from scipy.spatial import distance_matrix

    last = np.asarray(x_position)[-1]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        contact_map = (distance_matrix(last, last) < threshold).astype("int8")
    sparse = coo_matrix(contact_map)
    edge_index = np.array([sparse.row, sparse.col]).astype(np.int64)
    edge_attr = np.concatenate([last[edge_index[0]], last[edge_index[1]]], axis=1).astype(np.float32)
    return edge_index, edge_attr


def radius_graph_batched(pos: np.ndarray, threshold: float):
    """Frames [F,N,3] -> one PyG-style batch: per-frame construct_pairdata graphs concatenated with
    node ids offset by f*N (PairData.__inc__, dataset.py:41-45)."""
    eis, eas = [], []
    N = pos.shape[1]
    for f in range(pos.shape[0]):
        ei, ea = construct_pairdata_arrays(pos[f][None], threshold)
        eis.append(ei + f * N)
        eas.append(ea)
    return np.concatenate(eis, axis=1), np.concatenate(eas, axis=0)


def radius_graph_kdtree(pos: np.ndarray, threshold: float):
    """Scalable equivalent for large N (the dense path needs N^2 float64): cKDTree candidates with
    a slightly inflated radius, re-filtered with the reference's exact float64 predicate and put
    in (row, col) order.  Validated against the dense path in tests at N <= 2000."""
    from scipy.spatial import cKDTree

    p32 = np.asarray(pos, dtype=np.float32)
    p = p32.astype(np.float64)
    tree = cKDTree(p)
    pairs = tree.query_pairs(threshold * (1 + 1e-9) + 1e-12, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1], np.arange(len(p))])
    j = np.concatenate([pairs[:, 1], pairs[:, 0], np.arange(len(p))])
    d = p[j] - p[i]
    sq = d * d
    dist = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])
    keep = dist < threshold
    i, j = i[keep], j[keep]
    order = np.lexsort((j, i))
    edge_index = np.stack([i[order], j[order]]).astype(np.int64)
    edge_attr = np.concatenate([p32[edge_index[0]], p32[edge_index[1]]], axis=1)
    return edge_index, edge_attr


# --------------------------------------------------------------------------------------------
# Model-independent parameter helpers (deterministic synthetic weights for tests)
# --------------------------------------------------------------------------------------------
def make_net_params(ker_in: int, K: int, w: int, seed: int = 0, dtype=torch.float64) -> dict:
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights for DenseNet([ker_in,K,K,w*w])."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return ((torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype), (
            (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * b
        ).to(dtype)

    p = {}
    p["layers.0.weight"], p["layers.0.bias"] = lin(K, ker_in)
    p["layers.2.weight"], p["layers.2.bias"] = lin(K, K)
    p["layers.4.weight"], p["layers.4.bias"] = lin(w * w, K)
    return p
