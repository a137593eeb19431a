"""MPNN Q-network with the reference's module tree and checkpoint format (reference src/networks/mpnn.py:5-159;
state_dict keys / shapes in SURVEY.md appendix A.3), so the shipped `.pth` files load unchanged.

Two execution paths:
  * `forward(obs)` -- differentiable PyTorch ops on any device: the fp32 torch reference of the CUDA kernels (tests) and
    the autograd route of the DQN update when the user supplies a loss callable.  The edge stage uses the factorised form
    ReLU(W_e [a_ij ; x_j]) = ReLU(a_ij w0 + W_x x_j), i.e. two N x N contractions for {-1,0,1} couplings instead of
    the reference's [B,N,N,63] intermediate (mpnn.py:89-100); real-valued couplings take the dense route.
  * `engine_weights()` -- the same parameters as device pointers for the hand-written kernels
    (eco_dqn_b200.engine.BatchedSpinSystem.q_values / rollout, and eco_mpnn_grad for the DQN update's loss and
    gradients): that is the hot path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class EdgeAndNodeEmbeddingLayer(nn.Module):
    def __init__(self, n_obs_in, n_features):
        super().__init__()
        self.n_obs_in = n_obs_in
        self.n_features = n_features
        self.edge_embedding_NN = nn.Linear(int(n_obs_in + 1), n_features - 1, bias=False)
        self.edge_feature_NN = nn.Linear(n_features, n_features, bias=False)

    def forward(self, node_features, adj, norm):
        w = self.edge_embedding_NN.weight                      # (F-1, 1 + n_obs): column 0 multiplies a_ij
        proj = F.linear(node_features, w[:, 1:])               # P_j = W_x x_j             [B, N, F-1]
        w0 = w[:, 0]
        if bool(((adj == 0) | (adj == 1) | (adj == -1)).all()):
            pos, neg = (adj > 0).to(proj.dtype), (adj < 0).to(proj.dtype)
            summed = torch.matmul(pos, F.relu(proj + w0)) + torch.matmul(neg, F.relu(proj - w0))
        else:   # general weights: sum_j [a_ij != 0] ReLU(a_ij w0 + P_j), dense
            mask = (adj != 0).to(proj.dtype).unsqueeze(-1)
            summed = (F.relu(adj.unsqueeze(-1) * w0 + proj.unsqueeze(1)) * mask).sum(dim=2)
        embedded_edges = summed / norm
        return F.relu(self.edge_feature_NN(torch.cat([embedded_edges, norm / norm.max()], dim=-1)))   # mpnn.py:102


class UpdateNodeEmbeddingLayer(nn.Module):
    def __init__(self, n_features):
        super().__init__()
        self.message_layer = nn.Linear(2 * n_features, n_features, bias=False)
        self.update_layer = nn.Linear(2 * n_features, n_features, bias=False)

    def forward(self, current_node_embeddings, edge_embeddings, norm, adj):
        aggregated = torch.matmul(adj, current_node_embeddings) / norm                               # mpnn.py:115
        message = F.relu(self.message_layer(torch.cat([aggregated, edge_embeddings], dim=-1)))
        return F.relu(self.update_layer(torch.cat([current_node_embeddings, message], dim=-1)))


class ReadoutLayer(nn.Module):
    def __init__(self, n_features, n_hid=[], bias_pool=False, bias_readout=True):
        super().__init__()
        self.layer_pooled = nn.Linear(int(n_features), int(n_features), bias=bias_pool)
        if type(n_hid) != list:
            n_hid = [n_hid]
        sizes = [2 * n_features] + n_hid + [1]
        self.layers_readout = nn.ModuleList([nn.Linear(a, b, bias=bias_readout) for a, b in zip(sizes, sizes[1:])])

    def forward(self, node_embeddings):
        pooled = self.layer_pooled(node_embeddings.sum(dim=1) / node_embeddings.shape[1])            # mpnn.py:147
        features = F.relu(torch.cat([pooled.unsqueeze(1).expand_as(node_embeddings), node_embeddings], dim=-1))
        for i, layer in enumerate(self.layers_readout):
            features = layer(features)
            if i < len(self.layers_readout) - 1:
                features = F.relu(features)
        return features


class MPNN(nn.Module):
    def __init__(self, n_obs_in=7, n_layers=3, n_features=64, tied_weights=False, n_hid_readout=[]):
        super().__init__()
        self.n_obs_in = n_obs_in
        self.n_layers = n_layers
        self.n_features = n_features
        self.tied_weights = tied_weights
        self.node_init_embedding_layer = nn.Sequential(nn.Linear(n_obs_in, n_features, bias=False), nn.ReLU())
        self.edge_embedding_layer = EdgeAndNodeEmbeddingLayer(n_obs_in, n_features)
        if self.tied_weights:
            self.update_node_embedding_layer = UpdateNodeEmbeddingLayer(n_features)
        else:
            self.update_node_embedding_layer = nn.ModuleList(
                [UpdateNodeEmbeddingLayer(n_features) for _ in range(self.n_layers)])
        self.readout_layer = ReadoutLayer(n_features, n_hid_readout)
        self._engine_cache = None

    @torch.no_grad()
    def get_normalisation(self, adj):
        norm = torch.sum((adj != 0), dim=1).unsqueeze(-1)                                            # mpnn.py:34-38
        norm[norm == 0] = 1
        return norm.float()

    def forward(self, obs):
        """obs [B, n_obs_in + N, N] (or [n_obs_in + N, N]) -> Q [B, N]; like the reference the result is squeezed,
        so B == 1 gives [N] (mpnn.py:75).  Unlike the reference the caller's tensor is not transposed in place."""
        if obs.dim() == 2:
            obs = obs.unsqueeze(0)
        obs = obs.transpose(-1, -2)
        node_features = obs[:, :, 0:self.n_obs_in]
        adj = obs[:, :, self.n_obs_in:]
        norm = self.get_normalisation(adj)
        h = self.node_init_embedding_layer(node_features)
        e = self.edge_embedding_layer(node_features, adj, norm)
        for i in range(self.n_layers):
            layer = self.update_node_embedding_layer if self.tied_weights else self.update_node_embedding_layer[i]
            h = layer(h, e, norm, adj)
        return self.readout_layer(h).squeeze()

    # ------------------------------------------------------------------ bridge to the CUDA kernels
    def engine_weights(self, device=None):
        """Device-pointer view of the parameters for the CUDA kernels (re-packed only when a parameter changed)."""
        from .. import engine
        if self.n_obs_in not in (1, 7) or self.n_layers != 3 or self.n_features != 64 or self.tied_weights or \
                len(self.readout_layer.layers_readout) != 1:
            raise NotImplementedError("the CUDA kernels implement the reference configurations: n_obs_in=7 (ECO-DQN) or "
                                      "1 (S2V-DQN), 3 untied layers, 64 features, no hidden readout layer")
        version = tuple(p._version for p in self.parameters())
        ptrs = tuple(p.data_ptr() for p in self.parameters())
        cache = self._engine_cache
        if cache is not None and cache[1] == ptrs and cache[2].aliases:
            if cache[0] != version:            # same storage, new values (optimizer step): only the packed copy is stale
                cache[2].repack()
                self._engine_cache = (version, ptrs, cache[2])
        elif cache is None or cache[0] != version or cache[1] != ptrs:
            self._engine_cache = (version, ptrs, engine.MPNNWeights(self.state_dict(), device=device))
        return self._engine_cache[2]
