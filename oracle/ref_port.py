"""CPU restatement ("port") of the reference's inductive graph-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product path (inductive-recommendation_b200/)
never does and fails loudly without its CUDA library.

Parity status: PINNED against the reference itself -- tests/test_oracle_golden.py checks every
function here against tests/golden/*.npz, which oracle/make_golden.py produced by running the
unmodified /root/reference/{dataset,utils,model,trainer}.py on CPU (the reference ships no tests
or golden vectors of its own).  Third-party arithmetic the reference delegates
to DGL (unpinned ">= 0.8", README.md:8; gspmm 'mul','sum') is restated as a CSR SpMM.

Each function cites the reference lines it follows.  fp32 throughout, like the reference.
This port is also the CPU baseline that bench.py times (torch CSR / MKL, all host threads), so it
keeps the reference's cost structure: full propagation on every training batch and on every
evaluation batch, autograd backward, dense torch.optim.Adam, per-batch Python exclusion lists.
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- graph construction
def pairs_from_csr(indptr, items):
    """train_array = [[u, i] ...] user-major (dataset.py:149-151)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    users = np.repeat(np.arange(len(indptr) - 1, dtype=np.int64), np.diff(indptr))
    return users, np.asarray(items, dtype=np.int64)


def adjacency(n_users, n_items, users, items):
    """utils.py:42-50 generate_daj_mat: symmetric bipartite adjacency, duplicates summed, fp32 CSR."""
    n = n_users + n_items
    row = np.concatenate([users, items + n_users])
    col = np.concatenate([items + n_users, users])
    return sp.coo_matrix((np.ones(row.shape), (row, col)), shape=(n, n), dtype=np.float32).tocsr()


def norm_adjacency(n_users, n_items, users, items):
    """model.py:89-98 LightGCN.generate_graph: D^-1/2 A D^-1/2 with deg = max(1, rowsum), fp32.
    Returns a scipy CSR whose (indptr, indices, data) equal the reference's coalesced COO order."""
    adj = adjacency(n_users, n_items, users, items)
    degree = np.array(np.sum(adj, axis=1)).squeeze()
    degree = np.maximum(1., degree)
    d_inv = np.power(degree, -0.5)
    d_mat = sp.diags(d_inv, format='csr', dtype=np.float32)
    out = d_mat.dot(adj).dot(d_mat).tocsr()
    out.sort_indices()
    return out


def norm_adjacency_from_csr(n_users, n_items, indptr, items):
    """norm_adjacency for a user-major, per-user sorted, duplicate-free pair list given as CSR (the 100M-edge bench graph):
    the same matrix -- structure from the user-item CSR and its transpose instead of a 200M-entry COO sort, values
    fl(fl(d_r * 1) * d_c) formed elementwise in float32, which is what d_mat.dot(adj).dot(d_mat) computes entry by entry.
    tests/test_oracle_golden.py checks it bit for bit against norm_adjacency."""
    indptr = np.asarray(indptr, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    r = sp.csr_matrix((np.ones(items.size, dtype=np.float32), items, indptr), shape=(n_users, n_items))
    rt = r.T.tocsr()
    rt.sort_indices()
    n = n_users + n_items
    ptr = np.concatenate([indptr, indptr[-1] + rt.indptr[1:].astype(np.int64)])
    idx = np.concatenate([items + n_users, rt.indices.astype(np.int64)])
    degree = np.maximum(1., np.diff(ptr).astype(np.float32))
    d_inv = np.power(degree, -0.5)
    data = np.multiply(np.multiply(np.repeat(d_inv, np.diff(ptr)), np.float32(1.)), d_inv[idx])
    return sp.csr_matrix((data, idx, ptr), shape=(n, n))


def torch_csr(mat):
    mat = mat.tocsr()
    return torch.sparse_csr_tensor(torch.from_numpy(mat.indptr.astype(np.int64)),
                                   torch.from_numpy(mat.indices.astype(np.int64)),
                                   torch.from_numpy(mat.data.astype(np.float32)), size=mat.shape)


class _SpMM(torch.autograd.Function):
    """dgl.ops.gspmm(g,'mul','sum',X,vals) == A @ X with grad to X only (model.py:106)."""

    @staticmethod
    def forward(ctx, x, a, at):
        ctx.at = at
        return torch.sparse.mm(a, x)

    @staticmethod
    def backward(ctx, gy):
        return torch.sparse.mm(ctx.at, gy.contiguous()), None, None


def spmm(a, at, x):
    return _SpMM.apply(x, a, at)


# ----------------------------------------------------------------------------- template features
def build_feat(n_users, n_items, users, items, user_map, item_map):
    """model.py:4139-4175 IGCN.generate_feat (after the template maps are chosen): rows are nodes,
    columns are template users | template items | 2 shared dummy columns.  Returns scipy CSR of ones
    (duplicates summed) and row_sum (fp32)."""
    tu, ti = len(user_map), len(item_map)
    umap = np.full(n_users, -1, dtype=np.int64)
    for k, v in user_map.items():
        umap[int(k)] = v
    imap = np.full(n_items, -1, dtype=np.int64)
    for k, v in item_map.items():
        imap[int(k)] = v
    m_i = imap[items] >= 0
    m_u = umap[users] >= 0
    rows = np.concatenate([users[m_i], n_users + items[m_u], np.arange(n_users), n_users + np.arange(n_items)])
    cols = np.concatenate([tu + imap[items[m_i]], umap[users[m_u]],
                           np.full(n_users, tu + ti), np.full(n_items, tu + ti + 1)])
    feat = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n_users + n_items, tu + ti + 2),
                         dtype=np.float32).tocsr()
    feat.sort_indices()
    row_sum = np.array(np.sum(feat, axis=1)).squeeze().astype(np.float32)
    return feat, row_sum


def rank_nodes_sort(n_users, n_items, users, items):
    """utils.py:186-215 graph_rank_nodes(ranking_metric='sort'): column mass of the l1-row-normalised
    adjacency, descending argsort (np.argsort(...)[::-1])."""
    from sklearn.preprocessing import normalize
    adj = adjacency(n_users, n_items, users, items)
    nadj = normalize(adj, axis=1, norm='l1')
    um = np.array(np.sum(nadj[:, :n_users], axis=0)).squeeze()
    im = np.array(np.sum(nadj[:, n_users:], axis=0)).squeeze()
    return np.argsort(um)[::-1].copy(), np.argsort(im)[::-1].copy()


def template_maps(n_users, n_items, users, items, feature_ratio):
    """model.py:4140-4155: template (core) users/items -> column index maps."""
    if feature_ratio < 1.:
        ru, ri = rank_nodes_sort(n_users, n_items, users, items)
        cu = ru[:int(n_users * feature_ratio)]
        ci = ri[:int(n_items * feature_ratio)]
    else:
        cu = np.arange(n_users, dtype=np.int64)
        ci = np.arange(n_items, dtype=np.int64)
    return {int(u): k for k, u in enumerate(cu)}, {int(i): k for k, i in enumerate(ci)}


def feat_values(row_sum, feat_rows, alpha):
    """model.py:4127-4130 update_feat_mat: val[e] = row_sum[row[e]] ** ((alpha-1)/2 - 0.5) (torch.pow fp32)."""
    rs = torch.from_numpy(np.asarray(row_sum, dtype=np.float32))
    return torch.pow(rs[torch.from_numpy(np.asarray(feat_rows, dtype=np.int64))], (alpha - 1.) / 2. - 0.5).numpy()


def dropout_keep(nnz, p, generator=None):
    """model.py:4016-4021 NGCF.dropout_sp_mat: keep[e] = floor(1 - p + rand[e]) drawn on the CPU generator."""
    rnd = torch.rand(nnz, generator=generator)
    return torch.floor((1 - p) + rnd).type(torch.bool)


# ----------------------------------------------------------------------------- models
class LightGCNPort(torch.nn.Module):
    """model.py:79-127."""

    def __init__(self, n_users, n_items, users, items, emb0, n_layers, adj_sp=None):
        super().__init__()
        self.n_users, self.n_items, self.n_layers = n_users, n_items, n_layers
        self.adj_sp = adj_sp if adj_sp is not None else norm_adjacency(n_users, n_items, users, items)
        self.a = torch_csr(self.adj_sp)          # bit-wise symmetric: a^T == a
        self.embedding = torch.nn.Embedding(n_users + n_items, emb0.shape[1])
        with torch.no_grad():
            self.embedding.weight.copy_(torch.as_tensor(emb0))

    def layer0(self):
        return self.embedding.weight

    def get_rep(self):  # model.py:100-110
        x = self.layer0()
        reps = [x]
        for _ in range(self.n_layers):
            x = spmm(self.a, self.a, x)
            reps.append(x)
        return torch.stack(reps, dim=0).mean(dim=0)

    def bpr_forward(self, users, pos, neg):  # model.py:112-120
        rep = self.get_rep()
        e = self.embedding
        ue, pe, ne = e(users), e(self.n_users + pos), e(self.n_users + neg)
        l2 = torch.norm(ue, p=2, dim=1) ** 2 + torch.norm(pe, p=2, dim=1) ** 2 + torch.norm(ne, p=2, dim=1) ** 2
        return rep[users, :], rep[self.n_users + pos, :], rep[self.n_users + neg, :], l2

    def predict(self, users):  # model.py:122-127
        rep = self.get_rep()
        return torch.mm(rep[users, :], rep[self.n_users:, :].t())


def infonce(query, positive_key, temperature=0.1):
    """The `info_nce` package as SGL / HALF call it (model.py:206-214): InfoNCE(negative_mode='unpaired')(q, k, k) --
    rows L2-normalised (F.normalize), logits [q.k_i (positive) | q.K^T] / T, cross-entropy against index 0, mean."""
    q = F.normalize(query, dim=-1)
    k = F.normalize(positive_key, dim=-1)
    logits = torch.cat([torch.sum(q * k, dim=1, keepdim=True), q @ k.t()], dim=1) / temperature
    return F.cross_entropy(logits, torch.zeros(len(logits), dtype=torch.long), reduction='mean')


class SGLPort(LightGCNPort):
    """SGL (model.py:130-243) / HALF (:246-365): LightGCN propagation on the full graph plus one or two edge-dropped views
    (utils.py:91-103: a uniform sample of int(E * aug_rate) train pairs, degrees recomputed), contrasted with InfoNCE at
    the batch users.  `views`: list of (users, items) pair arrays of the kept edges (two for SGL, one for HALF)."""

    def __init__(self, n_users, n_items, users, items, emb0, n_layers, views, half=False):
        super().__init__(n_users, n_items, users, items, emb0, n_layers)
        self.half = half
        self.views = [torch_csr(norm_adjacency(n_users, n_items, vu, vi)) for vu, vi in views]
        assert len(self.views) == (1 if half else 2)

    def get_aug_rep(self, a):  # model.py:189-200
        x = self.layer0()
        reps = [x]
        for _ in range(self.n_layers):
            x = spmm(a, a, x)
            reps.append(x)
        return torch.stack(reps, dim=0).mean(dim=0)

    def bpr_forward(self, users, pos, neg):  # model.py:216-229 / :334-349: l2 on the FINAL representations
        rep = self.get_rep()
        ur, pr, nr = rep[users, :], rep[self.n_users + pos, :], rep[self.n_users + neg, :]
        l2 = torch.norm(ur, p=2, dim=1) ** 2 + torch.norm(pr, p=2, dim=1) ** 2 + torch.norm(nr, p=2, dim=1) ** 2
        a1 = self.get_aug_rep(self.views[0])[users, :]
        con = infonce(ur, a1) if self.half else infonce(a1, self.get_aug_rep(self.views[1])[users, :])
        return ur, pr, nr, l2, con


def contrastive_train_step(model, opt, batch, l2_reg, contrastive_reg):
    """One iteration of SGLTrainer / HALFTrainer.train_one_epoch (trainer.py:440-456)."""
    users, pos, neg = batch[:, 0], batch[:, 1], batch[:, 2]
    ur, pr, nr, l2, con = model.bpr_forward(users, pos, neg)
    bpr = F.softplus(torch.sum(ur * nr, dim=1) - torch.sum(ur * pr, dim=1)).mean()
    loss = bpr + l2_reg * l2.mean() + contrastive_reg * con.mean()
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.item(), float(con.detach())


class MFPort(torch.nn.Module):
    """model.py:56-76."""

    def __init__(self, user_emb0, item_emb0):
        super().__init__()
        self.n_users, self.n_items = user_emb0.shape[0], item_emb0.shape[0]
        self.user_embedding = torch.nn.Embedding(*user_emb0.shape)
        self.item_embedding = torch.nn.Embedding(*item_emb0.shape)
        with torch.no_grad():
            self.user_embedding.weight.copy_(torch.as_tensor(user_emb0))
            self.item_embedding.weight.copy_(torch.as_tensor(item_emb0))

    def bpr_forward(self, users, pos, neg):
        ue, pe, ne = self.user_embedding(users), self.item_embedding(pos), self.item_embedding(neg)
        l2 = torch.norm(ue, p=2, dim=1) ** 2 + torch.norm(pe, p=2, dim=1) ** 2 + torch.norm(ne, p=2, dim=1) ** 2
        return ue, pe, ne, l2

    def predict(self, users):
        return torch.mm(self.user_embedding(users), self.item_embedding.weight.t())


class IGCNPort(LightGCNPort):
    """model.py:4107-4220 (IGCN) and :4290-4297 (IMF = n_layers 0)."""

    def __init__(self, n_users, n_items, users, items, emb0, n_layers, dropout, user_map, item_map, w0=None,
                 alpha=1.0, delta=0.99):
        torch.nn.Module.__init__(self)
        self.n_users, self.n_items, self.n_layers = n_users, n_items, n_layers
        self.dropout, self.alpha, self.delta = dropout, alpha, delta
        self.user_map, self.item_map = user_map, item_map
        self.adj_sp = norm_adjacency(n_users, n_items, users, items)
        self.a = torch_csr(self.adj_sp)
        self.feat_sp, self.row_sum = build_feat(n_users, n_items, users, items, user_map, item_map)
        coo = self.feat_sp.tocoo()
        self.feat_rows, self.feat_cols = coo.row.astype(np.int64), coo.col.astype(np.int64)
        self.embedding = torch.nn.Embedding(self.feat_sp.shape[1], emb0.shape[1])
        self.w = torch.nn.Parameter(torch.ones(emb0.shape[1]) if w0 is None else torch.as_tensor(w0).clone())
        with torch.no_grad():
            self.embedding.weight.copy_(torch.as_tensor(emb0))
        self.forced_keep = None  # tests inject the reference's recorded dropout draw

    def feat_vals(self):
        return feat_values(self.row_sum, self.feat_rows, self.alpha)

    def feat_mat_anneal(self):  # model.py:4132-4134
        self.alpha *= self.delta

    def layer0(self):  # model.py:4189-4190 dropout_sp_mat + inductive_rep_layer (:4177-4186)
        vals = self.feat_vals()
        rows, cols = self.feat_rows, self.feat_cols
        if self.training:
            keep = self.forced_keep if self.forced_keep is not None else dropout_keep(len(vals), self.dropout)
            keep = np.asarray(keep, dtype=bool)
            rows, cols = rows[keep], cols[keep]
            vals = (torch.from_numpy(vals[keep]) / (1. - self.dropout)).numpy()
        f = sp.coo_matrix((vals, (rows, cols)), shape=self.feat_sp.shape, dtype=np.float32).tocsr()
        return spmm(torch_csr(f), torch_csr(f.T.tocsr()), self.embedding.weight)

    def bpr_forward(self, users, pos, neg):  # model.py:4046-4052 (NGCF.bpr_forward): L2 on gathered reps
        rep = self.get_rep()
        ur, pr, nr = rep[users, :], rep[self.n_users + pos, :], rep[self.n_users + neg, :]
        l2 = torch.norm(ur, p=2, dim=1) ** 2 + torch.norm(pr, p=2, dim=1) ** 2 + torch.norm(nr, p=2, dim=1) ** 2
        return ur, pr, nr, l2


# ----------------------------------------------------------------------------- training step
def bpr_loss(model, batch, l2_reg):
    """trainer.py:414-424: softplus(neg - pos).mean() + l2_reg * l2_norm_sq.mean()."""
    users, pos, neg = batch[:, 0], batch[:, 1], batch[:, 2]
    ur, pr, nr, l2 = model.bpr_forward(users, pos, neg)
    ps, ns = torch.sum(ur * pr, dim=1), torch.sum(ur * nr, dim=1)
    bpr = F.softplus(ns - ps).mean()
    return bpr, bpr + l2_reg * l2.mean(), (ur, pr, nr, l2)


def aux_loss(model, aux_batch):
    """trainer.py:542-549: BPR on raw template embeddings weighted by model.w."""
    au, ap, an = aux_batch[:, 0], aux_batch[:, 1], aux_batch[:, 2]
    tu = len(model.user_map)
    e = model.embedding
    eu, ep, en = e(au), e(ap + tu), e(an + tu)
    ps = torch.sum(eu * ep * model.w[None, :], dim=1)
    ns = torch.sum(eu * en * model.w[None, :], dim=1)
    return F.softplus(ns - ps).mean()


def train_step(model, opt, batch, l2_reg, aux_batch=None, aux_reg=0.0):
    """One iteration of BPRTrainer / IGCNTrainer.train_one_epoch (trainer.py:412-429, :531-558)."""
    _, loss, _ = bpr_loss(model, batch, l2_reg)
    if aux_batch is not None:
        loss = loss + aux_reg * aux_loss(model, aux_batch)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.item()


# ----------------------------------------------------------------------------- evaluation
def calculate_metrics(eval_data, rec_items, topks):
    """trainer.py:115-144, vectorised per user but the same arithmetic (fp32 hit matrix, log2(j+2))."""
    rec_items = np.asarray(rec_items)
    hit = np.zeros_like(rec_items, dtype=np.float32)
    for u in range(rec_items.shape[0]):
        if len(eval_data[u]):
            hit[u] = np.isin(rec_items[u], np.asarray(eval_data[u])).astype(np.float32)
    lens = np.array([len(x) for x in eval_data], dtype=np.int32)
    res = {'Precision': {}, 'Recall': {}, 'NDCG': {}}
    for k in topks:
        hit_num = np.sum(hit[:, :k], axis=1)
        precisions = hit_num / k
        with np.errstate(invalid='ignore', divide='ignore'):
            recalls = hit_num / lens
        max_hit = np.minimum(lens, k)
        max_hit_matrix = (np.arange(k)[None, :] < max_hit[:, None]).astype(np.float32)
        denom = np.log2(np.arange(2, k + 2, dtype=np.float32))[None, :]
        dcgs = np.sum(hit[:, :k] / denom, axis=1)
        idcgs = np.sum(max_hit_matrix / denom, axis=1)
        with np.errstate(invalid='ignore', divide='ignore'):
            ndcgs = dcgs / idcgs
        mask = max_hit > 0
        res['Precision'][k] = precisions[mask].mean()
        res['Recall'][k] = recalls[mask].mean()
        res['NDCG'][k] = ndcgs[mask].mean()
    return res


def topk_tiebreak(scores, k):
    """torch.topk leaves tie order unspecified (trainer.py:169); the stated contract is
    (score descending, item id ascending) -- what a stable sort on -score gives.  Selected per row with a partition
    (everything at or above the k-th best value, ties included) and only those few are sorted, so a 1M-item row
    costs O(n) instead of a full sort."""
    s = np.asarray(scores)
    n = s.shape[1]
    k = min(k, n)
    order = np.empty((s.shape[0], k), dtype=np.int64)
    for r in range(s.shape[0]):
        neg = -s[r]
        kth = np.partition(neg, k - 1)[k - 1]
        cand = np.nonzero(neg <= kth)[0] if not np.isnan(kth) else np.arange(n)   # ascending ids
        order[r] = cand[np.argsort(neg[cand], kind='stable')[:k]]
    return order, np.take_along_axis(s, order, axis=1)


def evaluate(model, train_lists, val_lists, eval_lists, split, topks, test_batch_size=512, banned=None,
             n_users=None):
    """trainer.py:146-170: per user batch predict (re-propagates!), mask, top-K.  Returns rec_items, metrics."""
    model.eval()
    n_users = model.n_users if n_users is None else n_users
    kmax = max(topks)
    rec = []
    with torch.no_grad():
        for s in range(0, n_users, test_batch_size):
            users = torch.arange(s, min(n_users, s + test_batch_size), dtype=torch.int64)
            scores = model.predict(users)
            if split != 'train':
                ex_u, ex_i = [], []
                for ui, u in enumerate(users.tolist()):
                    items = train_lists[u]
                    if split == 'test':
                        items = items + val_lists[u]
                    ex_u.extend([ui] * len(items))
                    ex_i.extend(items)
                scores[ex_u, ex_i] = -np.inf
            if banned is not None:
                scores[:, banned[0]:banned[1]] = -np.inf
            ids, _ = topk_tiebreak(scores.numpy(), kmax)
            rec.append(ids)
    rec = np.concatenate(rec, axis=0)
    return rec, calculate_metrics(eval_lists, rec, topks)
