"""Oracle: MPNN Q-network forward, fp32, CPU torch, arithmetic as the reference writes it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates reference src/networks/mpnn.py:

  MPNN.forward                      mpnn.py:40-77    -> mpnn_forward
  MPNN.get_normalisation            mpnn.py:34-38    -> degree_norm
  EdgeAndNodeEmbeddingLayer.forward mpnn.py:89-104   -> edge_embedding   (dense [B,N,N,8] -> [B,N,N,63], as written)
  UpdateNodeEmbeddingLayer.forward  mpnn.py:114-120  -> update_layer
  ReadoutLayer.forward              mpnn.py:143-159  -> readout

The reference's arithmetic library for this path is PyTorch ATen (pinned torch 1.12.1; 2.11 in this image), so
the restatement calls the same ATen ops in the same order on the same dtypes rather than re-deriving them.
Weights are passed as a dict keyed like the reference's state_dict, so the shipped
`.pth` checkpoints and the `w::*` arrays in tests/golden/*.npz load unchanged.
"""
import numpy as np
import torch
import torch.nn.functional as F

KEYS = (
    "node_init_embedding_layer.0.weight",                 # (64, 7)
    "edge_embedding_layer.edge_embedding_NN.weight",      # (63, 8)   col 0 multiplies a_ij
    "edge_embedding_layer.edge_feature_NN.weight",        # (64, 64)
    "update_node_embedding_layer.0.message_layer.weight",  # (64, 128)
    "update_node_embedding_layer.0.update_layer.weight",   # (64, 128)
    "update_node_embedding_layer.1.message_layer.weight",
    "update_node_embedding_layer.1.update_layer.weight",
    "update_node_embedding_layer.2.message_layer.weight",
    "update_node_embedding_layer.2.update_layer.weight",
    "readout_layer.layer_pooled.weight",                  # (64, 64)
    "readout_layer.layers_readout.0.weight",              # (1, 128)
    "readout_layer.layers_readout.0.bias",                # (1,)
)


def as_torch_weights(w):
    return {k: torch.as_tensor(np.asarray(w[k]), dtype=torch.float32) for k in KEYS}


def weights_from_npz(z):
    """Pull the `w::<state_dict key>` arrays out of a golden .npz."""
    return {k: np.asarray(z["w::" + k], dtype=np.float32) for k in KEYS}


def degree_norm(adj):
    norm = torch.sum((adj != 0), dim=1).unsqueeze(-1)      # mpnn.py:36 (adj symmetric: dim 1 == dim 2 count)
    norm[norm == 0] = 1
    return norm.float()


def edge_embedding(w, x, adj, norm):
    B, N, _ = adj.shape
    ef = torch.empty(B, N, N, 1 + x.shape[-1])
    ef[..., 0] = adj                                       # [a_ij ; x_j] at [b, i, j]   (mpnn.py:90-92)
    ef[..., 1:] = x.unsqueeze(1)
    ef.mul_((adj != 0).unsqueeze(-1))                      # mpnn.py:94
    emb = F.linear(ef.view(B, N * N, -1), w[KEYS[1]]).relu_()       # mpnn.py:96-97
    emb = emb.view(B, N, N, -1).sum(dim=2) / norm          # mpnn.py:98-100
    return F.relu(F.linear(torch.cat([emb, norm / norm.max()], dim=-1), w[KEYS[2]]))   # mpnn.py:102


def update_layer(w, layer, h, e, norm, adj):
    agg = torch.matmul(adj, h) / norm                      # mpnn.py:115
    msg = F.relu(F.linear(torch.cat([agg, e], dim=-1), w[KEYS[3 + 2 * layer]]))     # mpnn.py:117
    return F.relu(F.linear(torch.cat([h, msg], dim=-1), w[KEYS[4 + 2 * layer]]))    # mpnn.py:118


def readout(w, h):
    pooled = F.linear(h.sum(dim=1) / h.shape[1], w[KEYS[9]])           # mpnn.py:147
    f = F.relu(torch.cat([pooled.unsqueeze(1).expand_as(h), h], dim=-1))   # mpnn.py:148-150
    return F.linear(f, w[KEYS[10]], w[KEYS[11]])                        # mpnn.py:152-157


@torch.no_grad()
def mpnn_forward(weights, obs):
    """obs: float32 [B, n_obs+N, N] (rows 0..n_obs-1 features, then the adjacency) -> Q float32 [B, N]; n_obs is taken from
    the weights (7, or 1 for the S2V-DQN networks).

    Unlike the reference this does not transpose the caller's tensor in place (mpnn.py:44, quirk A.4-2)
    and always returns [B, N] (the reference squeezes B == 1 away, mpnn.py:75)."""
    w = weights if isinstance(next(iter(weights.values())), torch.Tensor) else as_torch_weights(weights)
    obs = torch.as_tensor(obs, dtype=torch.float32)
    if obs.dim() == 2:
        obs = obs.unsqueeze(0)
    obs = obs.transpose(-1, -2)
    n_obs = w[KEYS[0]].shape[1]                            # 7 (ECO-DQN) or 1 (S2V-DQN: the spin only)
    x = obs[:, :, :n_obs]
    adj = obs[:, :, n_obs:]
    norm = degree_norm(adj)
    h = F.relu(F.linear(x, w[KEYS[0]]))                    # mpnn.py:55
    e = edge_embedding(w, x, adj, norm)
    for layer in range(3):                                 # mpnn.py:68-72 (untied weights)
        h = update_layer(w, layer, h, e, norm, adj)
    return readout(w, h).squeeze(-1)


@torch.no_grad()
def mpnn_forward_blocked(weights, x, adj, rows_per_block=64, norm_max=None):
    """mpnn_forward for ONE large graph without the [N, N, 63] intermediate of mpnn.py:90-98: the edge stage is
    evaluated for `rows_per_block` target vertices i at a time with exactly the reference's per-row arithmetic
    (concatenate [a_ij ; x_j], mask, linear, ReLU, sum over j), everything else as written.  Pinned against
    mpnn_forward in tests/test_oracle_golden.py; used by the GPU parity tests at N = 1100 ... 2048 where the dense
    tensor would need gigabytes per episode.

    x: float32 [N, n_obs] vertex observations, adj: float32 [N, N].  Returns Q float32 [N]."""
    w = weights if isinstance(next(iter(weights.values())), torch.Tensor) else as_torch_weights(weights)
    x = torch.as_tensor(x, dtype=torch.float32)
    adj = torch.as_tensor(adj, dtype=torch.float32)
    N = adj.shape[0]
    norm = degree_norm(adj.unsqueeze(0))[0]                # [N, 1]
    h = F.relu(F.linear(x, w[KEYS[0]]))
    emb = torch.empty(N, w[KEYS[1]].shape[0])
    for i0 in range(0, N, rows_per_block):
        a = adj[i0:i0 + rows_per_block]                    # [R, N]
        ef = torch.empty(a.shape[0], N, 1 + x.shape[-1])
        ef[..., 0] = a
        ef[..., 1:] = x.unsqueeze(0)
        ef.mul_((a != 0).unsqueeze(-1))
        emb[i0:i0 + rows_per_block] = F.linear(ef, w[KEYS[1]]).relu_().sum(dim=1)
    emb = emb / norm
    nm = norm.max() if norm_max is None else torch.tensor(float(norm_max))
    e = F.relu(F.linear(torch.cat([emb, norm / nm], dim=-1), w[KEYS[2]]))
    h, e, norm, adj = h.unsqueeze(0), e.unsqueeze(0), norm.unsqueeze(0), adj.unsqueeze(0)
    for layer in range(3):
        h = update_layer(w, layer, h, e, norm, adj)
    return readout(w, h).squeeze(-1)[0]
