"""Drop-in for the reference's model.py, hot-path classes only: MF, LightGCN, IGCN, IMF, SGL, HALF.

Surface kept from /root/reference/model.py: get_model (:20-25), BasicModel (:35-53), MF (:56-76), LightGCN (:79-127),
IGCN (:4107-4220), IMF (:4290-4297) -- same constructor config keys, attributes (embedding, w, user_map, item_map,
norm_adj, feat_mat, alpha, row_sum), methods (get_rep, bpr_forward, predict, generate_graph, generate_feat,
update_feat_mat, feat_mat_anneal, save, load) and state_dict keys.  The arithmetic the reference hands to DGL / ATen
(gspmm, index gather/put, mm) runs in libb200rec's sm_100a kernels through b200rec.ops; there is no CPU path, models
must live on a CUDA device.  SGL / HALF (:130-365) are built on the same operators.  Out of scope
(SURVEY.md section 2.1): NGCF proper, DOSE_*, ItemKNN, Popularity,
MultiVAE, NeuMF, IMCGAE, IDCF.

Additions: `recommend()` (fused score + mask + top-K, used by trainer.eval instead of predict + topk), an eval-mode
cache of get_rep() (the reference re-propagates for every 512-user batch, model.py:123), `dropout_rng` config key
('device' = Philox mask drawn in a kernel, 'host' = the reference's CPU torch.rand draw, model.py:4020).
"""
import sys

import numpy as np
import torch
import torch.nn as nn
from torch.nn.init import normal_

from b200rec import graph as b2graph
from b200rec import ops
from utils import graph_rank_nodes


def get_model(config, dataset):
    config = config.copy()
    config['dataset'] = dataset
    cls = getattr(sys.modules[__name__], config['name'])
    return cls(config)


class BasicModel(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        self.config = model_config
        self.name = model_config['name']
        self.device = torch.device(model_config['device'])
        self.n_users = model_config['dataset'].n_users
        self.n_items = model_config['dataset'].n_items
        self.trainable = True
        self._rep_cache = None

    def predict(self, users):
        raise NotImplementedError

    def save(self, path):
        torch.save(self.state_dict(), path)

    def load(self, path):
        self.load_state_dict(torch.load(path, map_location=self.device))

    # ---- shared helpers ----
    def _cache_key(self):
        return None

    def _cached_rep(self, compute):
        """eval-mode memo of get_rep(): valid while no parameter / graph / alpha changed."""
        if self.training or torch.is_grad_enabled():
            return compute()
        key = self._cache_key()
        if self._rep_cache is None or self._rep_cache[0] != key:
            self._rep_cache = (key, compute())
        return self._rep_cache[1]

    def _all_columns(self, rep):
        """with the embedding dimension sharded over GPUs (b200rec.dist.DimShard) scoring needs every column"""
        shard = getattr(self, '_dim_shard', None)
        return rep if shard is None else shard.gather_cols(rep)

    def recommend(self, users, k, excl_a=None, excl_b=None, banned=None, precision=0):
        """Fused full-rank scoring + masking + top-K (trainer.py:152-169): returns (ids int32 [b,k], scores [b,k]),
        ordered (score desc, item id asc)."""
        with torch.no_grad():
            table_u, table_i = self.score_tables()
            return ops.score_topk(table_u, users.contiguous(), table_i, k, excl_a, excl_b, banned, precision)


class MF(BasicModel):
    def __init__(self, model_config):
        super().__init__(model_config)
        self.embedding_size = model_config['embedding_size']
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        normal_(self.user_embedding.weight, std=0.1)
        normal_(self.item_embedding.weight, std=0.1)
        self.to(device=self.device)
        # one contiguous [U+I, D] buffer behind both tables so the fused kernels see a single row space
        joint = torch.cat([self.user_embedding.weight.data, self.item_embedding.weight.data], dim=0).contiguous()
        self._joint = joint
        self.user_embedding.weight = nn.Parameter(joint[:self.n_users])
        self.item_embedding.weight = nn.Parameter(joint[self.n_users:])

    def get_rep(self):
        return self._joint

    def score_tables(self):
        joint = self._all_columns(self._joint)
        return joint[:self.n_users], joint[self.n_users:]

    def bpr_forward(self, users, pos_items, neg_items):
        ue = ops.gather_rows(self.user_embedding.weight, users)
        pe = ops.gather_rows(self.item_embedding.weight, pos_items)
        ne = ops.gather_rows(self.item_embedding.weight, neg_items)
        l2_norm_sq = (ue * ue).sum(1) + (pe * pe).sum(1) + (ne * ne).sum(1)
        return ue, pe, ne, l2_norm_sq

    def predict(self, users):
        with torch.no_grad():
            tu, ti = self.score_tables()
            return ops.score_dense(tu, users.contiguous(), ti)


class LightGCN(BasicModel):
    def __init__(self, model_config):
        super().__init__(model_config)
        self.embedding_size = model_config['embedding_size']
        self.n_layers = model_config['n_layers']
        self.embedding = nn.Embedding(self.n_users + self.n_items, self.embedding_size)
        self.norm_adj = self.generate_graph(model_config['dataset'])
        normal_(self.embedding.weight, std=0.1)
        self.to(device=self.device)

    def generate_graph(self, dataset):
        """D^-1/2 A D^-1/2 as a device CSR operand (b200rec.graph.CsrOperand; it also answers .indices()/.values()/
        .shape like the reference's sparse tensor)."""
        users, items = dataset.train_pairs()
        return b2graph.build_norm_adj(dataset.n_users, dataset.n_items, torch.from_numpy(np.ascontiguousarray(users)),
                                      torch.from_numpy(np.ascontiguousarray(items)), device=self.device,
                                      d=self.embedding_size)

    def layer0(self):
        return self.embedding.weight

    def _cache_key(self):
        return (self.embedding.weight._version, self.embedding.weight.data_ptr(), id(self.norm_adj))

    def get_rep(self):
        return self._cached_rep(lambda: ops.propagate(self.norm_adj, self.layer0(), self.n_layers))

    def score_tables(self):
        rep = self.get_rep()
        if getattr(self, '_full_src', None) is not rep:  # column all-gather once per cached representation
            self._full_src, self._full_rep = rep, self._all_columns(rep)
        rep = self._full_rep
        return rep, rep[self.n_users:]

    def bpr_forward(self, users, pos_items, neg_items):
        rep = self.get_rep()
        w = self.embedding.weight
        ue = ops.gather_rows(w, users)
        pe = ops.gather_rows(w, pos_items, self.n_users)
        ne = ops.gather_rows(w, neg_items, self.n_users)
        l2_norm_sq = (ue * ue).sum(1) + (pe * pe).sum(1) + (ne * ne).sum(1)
        users_r = ops.gather_rows(rep, users)
        pos_items_r = ops.gather_rows(rep, pos_items, self.n_users)
        neg_items_r = ops.gather_rows(rep, neg_items, self.n_users)
        return users_r, pos_items_r, neg_items_r, l2_norm_sq

    def predict(self, users):
        with torch.no_grad():
            rep, items = self.score_tables()
            return ops.score_dense(rep, users.contiguous(), items)


class SGL(LightGCN):
    """LightGCN + self-supervised contrast between two edge-dropped views (reference model.py:130-243).

    `norm_aug_adj1/2` are D^-1/2 A' D^-1/2 of a uniform sample of int(E * aug_rate) train edges each (utils.py:91-103),
    redrawn by update_aug_adj() at every epoch end.  bpr_forward returns the reference's 5-tuple; the contrastive term is
    InfoNCE (temperature 0.1 -- the reference never passes its `taugh` key on) between the views' rows of the batch users.
    The reference samples the kept edges with Python's `random.sample`; here the subset comes from a seeded device
    permutation (same distribution, different stream), or from `keep_index` when a caller supplies the subset."""
    n_views = 2

    def __init__(self, model_config):
        super().__init__(model_config)
        self.alpha = 1.
        self.delta = model_config.get('delta', 0.99)
        self.taugh = model_config.get('taugh', 0.2)
        self.aug_rate = model_config.get('aug_rate', 0.8)
        self.temperature = 0.1  # InfoNCE() default; see class docstring
        self.times = model_config.get('times', 0)
        self._aug_gen = torch.Generator(device=self.device)
        self._aug_gen.manual_seed(int(model_config.get('aug_seed', 2021)))
        self.aug_version = 0
        dataset = model_config['dataset']
        self.norm_aug_adj1 = self.generate_drop_graph(dataset)
        self.norm_aug_adj2 = self.generate_drop_graph(dataset) if self.n_views == 2 else None

    def _train_pairs_dev(self, dataset):
        if getattr(self, '_pairs_src', None) is not dataset:
            users, items = dataset.train_pairs()
            self._pairs = (torch.from_numpy(np.ascontiguousarray(users)).to(self.device),
                           torch.from_numpy(np.ascontiguousarray(items)).to(self.device))
            self._pairs_src = dataset
        return self._pairs

    def generate_drop_graph(self, dataset, keep_index=None):
        """normalised adjacency of an edge subset: exactly int(E * aug_rate) distinct train pairs (utils.py:93-94), degrees
        recomputed on the subset (model.py:166-176)"""
        users, items = self._train_pairs_dev(dataset)
        n_edges = users.numel()
        if keep_index is None:
            n_keep = int(n_edges * self.aug_rate)
            keep_index = torch.randperm(n_edges, generator=self._aug_gen, device=self.device)[:n_keep]
            keep_index = torch.sort(keep_index).values
        else:
            keep_index = torch.as_tensor(keep_index, dtype=torch.int64, device=self.device)
        return b2graph.build_norm_adj(dataset.n_users, dataset.n_items, users[keep_index], items[keep_index],
                                      device=self.device, d=self.embedding_size)

    def get_aug_rep(self, norm_aug_adj):
        return ops.propagate(norm_aug_adj, self.layer0(), self.n_layers)

    def cal_loss(self, users_r, aug_users_r):
        return ops.infonce(users_r, aug_users_r, self.temperature)

    def _views(self, rep, users):
        a1 = ops.gather_rows(self.get_aug_rep(self.norm_aug_adj1), users)
        a2 = ops.gather_rows(self.get_aug_rep(self.norm_aug_adj2), users)
        return a1, a2

    def bpr_forward(self, users, pos_items, neg_items):
        rep = self.get_rep()
        users_r = ops.gather_rows(rep, users)
        pos_items_r = ops.gather_rows(rep, pos_items, self.n_users)
        neg_items_r = ops.gather_rows(rep, neg_items, self.n_users)
        l2_norm_sq = (users_r * users_r).sum(1) + (pos_items_r * pos_items_r).sum(1) + (neg_items_r * neg_items_r).sum(1)
        q, k = self._views(users_r, users)
        return users_r, pos_items_r, neg_items_r, l2_norm_sq, self.cal_loss(q, k)

    def update_aug_adj(self):
        dataset = self.config['dataset']
        self.norm_aug_adj1 = self.generate_drop_graph(dataset)
        if self.n_views == 2:
            self.norm_aug_adj2 = self.generate_drop_graph(dataset)
        self.aug_version += 1


class HALF(SGL):
    """One augmented view contrasted with the full graph's representation (reference model.py:246-365)."""
    n_views = 1

    def _views(self, users_r, users):
        return users_r, ops.gather_rows(self.get_aug_rep(self.norm_aug_adj1), users)


class IGCN(BasicModel):
    def __init__(self, model_config):
        super().__init__(model_config)
        self.embedding_size = model_config['embedding_size']
        self.n_layers = model_config['n_layers']
        self.dropout = model_config['dropout']
        self.feature_ratio = model_config['feature_ratio']
        self.dropout_rng = model_config.get('dropout_rng', 'device')
        self.dropout_seed = model_config.get('dropout_seed', 2021)
        self._drop_step = None
        self.norm_adj = self.generate_graph(model_config['dataset'])
        self.alpha = 1.
        self.delta = model_config.get('delta', 0.99)
        self.feat_mat, self.user_map, self.item_map, self.row_sum = \
            self.generate_feat(model_config['dataset'], ranking_metric=model_config.get('ranking_metric', 'sort'))
        self.update_feat_mat()
        self.embedding = nn.Embedding(self.feat_mat.n_cols, self.embedding_size)
        self.w = nn.Parameter(torch.ones([self.embedding_size], dtype=torch.float32, device=self.device))
        normal_(self.embedding.weight, std=0.1)
        self.to(device=self.device)

    generate_graph = LightGCN.generate_graph

    def update_feat_mat(self):
        """F[r, :] = row_sum[r] ** ((alpha-1)/2 - 0.5): only the per-row scale vector changes."""
        self.row_scale = self.feat_mat.row_scale(self.alpha)
        self._rep_cache = None

    def feat_mat_anneal(self):
        self.alpha *= self.delta
        self.update_feat_mat()

    def generate_feat(self, dataset, is_updating=False, ranking_metric=None):
        """Template maps + feature operand.  is_updating=True keeps the stored maps, so nodes that joined after
        training (ids beyond the old ranges) are expressed through template columns only -- the inductive path."""
        n_users, n_items = dataset.n_users, dataset.n_items
        if not is_updating:
            if self.feature_ratio < 1.:
                ranked_users, ranked_items = graph_rank_nodes(dataset, ranking_metric)
                core_users = ranked_users[:int(self.n_users * self.feature_ratio)]
                core_items = ranked_items[:int(self.n_items * self.feature_ratio)]
            else:
                core_users = np.arange(self.n_users, dtype=np.int64)
                core_items = np.arange(self.n_items, dtype=np.int64)
            user_map = {int(u): k for k, u in enumerate(core_users)}
            item_map = {int(i): k for k, i in enumerate(core_items)}
        else:
            user_map, item_map = self.user_map, self.item_map
            self.n_users, self.n_items = n_users, n_items
        ut = np.full(n_users, -1, dtype=np.int64)
        it = np.full(n_items, -1, dtype=np.int64)
        ut[np.fromiter(user_map.keys(), np.int64, len(user_map))] = np.fromiter(user_map.values(), np.int64, len(user_map))
        it[np.fromiter(item_map.keys(), np.int64, len(item_map))] = np.fromiter(item_map.values(), np.int64, len(item_map))
        users, items = dataset.train_pairs()
        feat, _, _ = b2graph.build_feat(n_users, n_items, torch.from_numpy(np.ascontiguousarray(users)),
                                        torch.from_numpy(np.ascontiguousarray(items)), torch.from_numpy(ut),
                                        torch.from_numpy(it), device=self.device)
        return feat, user_map, item_map, feat.row_sum

    def _keep_bits(self):
        """edge-dropout mask for this call (training only), in the coalesced row-major edge order."""
        nnz = self.feat_mat.nnz
        if self.dropout_rng == 'host':
            rnd = torch.rand(nnz)  # CPU generator, exactly the reference's draw
            keep = torch.floor((1 - self.dropout) + rnd).type(torch.bool)
            return ops.pack_keep_bits(keep.to(self.device))
        if self._drop_step is None:
            self._drop_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        bits = ops.dropout_bits(nnz, self.dropout, self.dropout_seed, self._drop_step)
        ops.step_advance(self._drop_step)
        return bits

    def layer0(self):
        if self.training and self.dropout > 0.:
            return ops.inductive_layer(self.feat_mat, self.embedding.weight, self.row_scale, self._keep_bits(),
                                       1. / (1. - self.dropout))
        return ops.inductive_layer(self.feat_mat, self.embedding.weight, self.row_scale)

    def _cache_key(self):
        return (self.embedding.weight._version, self.embedding.weight.data_ptr(), id(self.norm_adj), id(self.feat_mat),
                self.alpha)

    def get_rep(self):
        return self._cached_rep(lambda: ops.propagate(self.norm_adj, self.layer0(), self.n_layers))

    score_tables = LightGCN.score_tables
    predict = LightGCN.predict

    def bpr_forward(self, users, pos_items, neg_items):
        rep = self.get_rep()
        users_r = ops.gather_rows(rep, users)
        pos_items_r = ops.gather_rows(rep, pos_items, self.n_users)
        neg_items_r = ops.gather_rows(rep, neg_items, self.n_users)
        l2_norm_sq = (users_r * users_r).sum(1) + (pos_items_r * pos_items_r).sum(1) + (neg_items_r * neg_items_r).sum(1)
        return users_r, pos_items_r, neg_items_r, l2_norm_sq

    def save(self, path):
        params = {'sate_dict': self.state_dict(), 'user_map': self.user_map, 'item_map': self.item_map,
                  'alpha': self.alpha}  # key spelling kept for checkpoint compatibility with the reference
        torch.save(params, path)

    def load(self, path):
        params = torch.load(path, map_location=self.device, weights_only=False)
        self.load_state_dict(params['sate_dict'])
        self.user_map = params['user_map']
        self.item_map = params['item_map']
        self.alpha = params['alpha']
        self.feat_mat, _, _, self.row_sum = self.generate_feat(self.config['dataset'], is_updating=True)
        self.update_feat_mat()


class IMF(IGCN):
    """IGCN without propagation: the inductive layer only (reference config sets n_layers = 0)."""

    def get_rep(self):
        return self._cached_rep(self.layer0)


class DOSE_aug(IGCN):
    """IGCN + a contrast between its representation and the one propagated on a graph EDITED by similarity mining
    (reference model.py:367-613, trainer DOSEaugTrainer trainer.py:255-303): `cal_cos_sim` picks the `aug_num`
    (user, item) pairs with the LOWEST cosine similarity (the item vectors are negated before the top-k, :507) and
    `generate_aug_graph` adds them to the train graph (utils.py:71-89: set union, so every entry stays 1).  bpr_forward
    returns the reference's 5-tuple; the contrastive term is InfoNCE (temperature 0.1: `taugh` is never passed on) between
    the batch users' rows of the two representations, each built from its OWN edge-dropout draw of the feature matrix
    (get_def_rep and get_aug_rep both call dropout_sp_mat, :472, :487).

    Mining runs on the evaluation kernels (b200rec.mining.lowest_cosine_pairs: per-row top-128 on L2-normalised rows +
    b200rec_topk_global) instead of a dense U x I sklearn matrix on the host.  Reference defects, kept or stated:
      * the second half of the mined pairs is mapped to (user, item) with the wrong offset (:537-540); config key
        `mining_offsets`: 'reference' (default, same pairs as the reference) | 'fixed';
      * `DOSE_aug.update_aug_adj` calls `self.generate_drop_graph`, which the class does not define (:571-575), so the
        reference raises AttributeError at the end of the first epoch; here it re-mines and rebuilds the augmented
        graph (what the commented-out line in DOSEaugTrainer, trainer.py:301, intended).
    `cal_cos_sim` returns an int64 [m, 2] device tensor (`.tolist()` gives the reference's list of [user, item])."""
    edit = 'add'

    def __init__(self, model_config):
        super().__init__(model_config)
        self.taugh = model_config.get('taugh', 0.2)
        self.aug_num = int(model_config['aug_num'])
        self.times = model_config.get('times', 0.1)
        self.temperature = 0.1
        self.mining_offsets = model_config.get('mining_offsets', 'reference')
        self.mining_precision = model_config.get('mining_precision', 0)
        self.aug_version = 0
        self.norm_aug_adj = self._edited_graph(model_config['dataset'])

    _train_pairs_dev = SGL._train_pairs_dev

    def get_def_rep(self):
        return self.get_rep()

    def get_aug_rep(self, norm_aug_adj):
        return ops.propagate(norm_aug_adj, self.layer0(), self.n_layers)

    def cal_cos_sim(self):
        from b200rec import mining
        with torch.no_grad():
            rep = self.get_def_rep()
            return mining.lowest_cosine_pairs(rep[:self.n_users], rep[self.n_users:], self.aug_num,
                                              reference_offsets=self.mining_offsets == 'reference',
                                              precision=self.mining_precision)

    cal_cos_sim_v2 = cal_cos_sim

    def _edited_graph(self, dataset, aug_idx=None):
        """train graph with the mined pairs added ('add': utils.py:71-89) or removed ('drop': utils.py:126-141), as a
        normalised adjacency with degrees recomputed on the edited graph (model.py:411-422)"""
        pairs = self.cal_cos_sim() if aug_idx is None else torch.as_tensor(np.asarray(aug_idx), dtype=torch.int64)
        pairs = pairs.to(self.device).reshape(-1, 2)
        users, items = self._train_pairs_dev(dataset)
        n_items = dataset.n_items
        train_keys = users * n_items + items
        mined_keys = pairs[:, 0] * n_items + pairs[:, 1]
        if self.edit == 'add':
            keys = torch.unique(torch.cat([train_keys, mined_keys]))
        else:
            keys = train_keys[~torch.isin(train_keys, mined_keys)]
        return b2graph.build_norm_adj(dataset.n_users, dataset.n_items, torch.div(keys, n_items, rounding_mode='floor'),
                                      keys % n_items, device=self.device, d=self.embedding_size)

    def generate_aug_graph(self, dataset, aug_idx=None):
        return self._edited_graph(dataset, aug_idx)

    def cal_loss(self, users_r, aug_users_r):
        return ops.infonce(users_r, aug_users_r, self.temperature)

    def bpr_forward(self, users, pos_items, neg_items):
        users_r, pos_items_r, neg_items_r, l2_norm_sq = super().bpr_forward(users, pos_items, neg_items)
        aug_users_r = ops.gather_rows(self.get_aug_rep(self.norm_aug_adj), users)
        return users_r, pos_items_r, neg_items_r, l2_norm_sq, self.cal_loss(users_r, aug_users_r)

    def update_aug_adj(self):
        self.norm_aug_adj = self._edited_graph(self.config['dataset'])
        self.aug_version += 1


class DOSE_drop3(DOSE_aug):
    """The same model with the mined pairs REMOVED from the train graph where they are train edges (reference
    model.py:2544-2863, `generate_drop_graph` -> utils.generate_drop_daj_mat3; trainer DOSEdropTrainer trainer.py:304-353)."""
    edit = 'drop'

    def generate_drop_graph(self, dataset, aug_idx=None):
        return self._edited_graph(dataset, aug_idx)



class DOSE_drop2(DOSE_aug):
    """IGCN contrasted with its propagation on a RANDOM edge subset of the train graph (reference model.py:1673-1962,
    used by the reference's Gowalla config): `generate_drop_graph` keeps int(E * aug_rate) train pairs drawn without
    replacement (utils.generate_drop_daj_mat, utils.py:91-103 -- random.sample there, a seeded device permutation here),
    degrees recomputed on the subset; `update_aug_adj` redraws it after every epoch (trainer.py:350-351).  No mining:
    `aug_num` is read like the reference does but only `cal_cos_sim` (never called by the trainer) uses it."""
    edit = 'random'

    def __init__(self, model_config):
        self.aug_rate = model_config.get('aug_rate', 0.2)
        self._aug_gen = None
        super().__init__(model_config)

    def _edited_graph(self, dataset, aug_idx=None):
        if self._aug_gen is None:
            self._aug_gen = torch.Generator(device=self.device)
            self._aug_gen.manual_seed(int(self.config.get('aug_seed', 2021)))
        return SGL.generate_drop_graph(self, dataset, keep_index=aug_idx)

    def generate_drop_graph(self, dataset, keep_index=None):
        return self._edited_graph(dataset, keep_index)


class DOSE_aug_drop2(DOSE_aug):
    """IGCN contrasted with its propagation on the train graph PLUS the `aug_num` most similar (user, item) pairs among the
    low-degree nodes (reference model.py:3182-3430): `cal_cos_sim` ranks users and items by train degree, leaves out the
    top `aug_ratio` fraction of each, and takes the flat top-k of the cosine matrix of what remains (:3290-3324).
    The reference builds two graphs from two separate minings -- `norm_aug_adj` (utils.generate_aug_daj_mat) and
    `norm_drop_adj` (utils.generate_drop_daj_mat2) -- and contrasts with the second one (:3391-3404).
    Reference defects, stated: `generate_drop_graph` calls generate_drop_daj_mat2 without its `aug_rate` argument
    (model.py:3241 vs utils.py:105), so the reference class raises TypeError when constructed; the function itself
    discards its `random.sample` (utils.py:110), i.e. it returns the same union graph as generate_aug_daj_mat.  Here both
    graphs are that union, each from its own mining pass like the reference's constructor."""
    edit = 'add'

    def __init__(self, model_config):
        self.aug_ratio = model_config.get('aug_ratio', 0.2)
        self._core = None
        super().__init__(model_config)
        self.norm_drop_adj = self._edited_graph(model_config['dataset'])

    def _core_nodes(self, dataset):
        if self._core is None:
            import utils as _utils
            ranked_users, ranked_items = _utils.graph_rank_nodes(dataset, 'degree')
            cu = ranked_users[int(self.n_users * self.aug_ratio):]
            ci = ranked_items[int(self.n_items * self.aug_ratio):]
            self._core = (torch.from_numpy(np.ascontiguousarray(cu)).to(self.device),
                          torch.from_numpy(np.ascontiguousarray(ci)).to(self.device))
        return self._core

    def cal_cos_sim(self, dataset=None):
        from b200rec import mining
        core_users, core_items = self._core_nodes(self.config['dataset'] if dataset is None else dataset)
        with torch.no_grad():
            rep = self.get_def_rep()
            u, i, _ = mining.pair_topk_global(rep[core_users].contiguous(), rep[self.n_users + core_items].contiguous(),
                                              self.aug_num, precision=self.mining_precision)
        return torch.stack([core_users[u], core_items[i]], dim=1)

    cal_cos_sim_v2 = cal_cos_sim

    def generate_drop_graph(self, dataset, aug_idx=None):
        return self._edited_graph(dataset, aug_idx)

    def get_drop_rep(self, norm_drop_adj):
        return self.get_aug_rep(norm_drop_adj)

    def bpr_forward(self, users, pos_items, neg_items):
        users_r, pos_items_r, neg_items_r, l2_norm_sq = IGCN.bpr_forward(self, users, pos_items, neg_items)
        aug_users_r = ops.gather_rows(self.get_drop_rep(self.norm_drop_adj), users)
        return users_r, pos_items_r, neg_items_r, l2_norm_sq, self.cal_loss(users_r, aug_users_r)

    def update_aug_adj(self):  # not defined by the reference class (DOSEdropTrainer would raise): both graphs are re-mined
        self.norm_aug_adj = self._edited_graph(self.config['dataset'])
        self.norm_drop_adj = self._edited_graph(self.config['dataset'])
        self.aug_version += 1
