"""Drop-in for the reference's trainer.py, hot-path classes only: BasicTrainer, BPRTrainer, IGCNTrainer, SGLTrainer,
HALFTrainer.

Surface kept from /root/reference/trainer.py: get_trainer (:16-22); BasicTrainer (:25-253) with train, eval,
calculate_metrics, inductive_eval, initialize_optimizer, record and the attributes topks / epoch / best_ndcg /
save_path / opt / dataloader / test_user_loader; BPRTrainer (:403-429); IGCNTrainer (:518-561).  Config keys are the
reference's (config.py); optional extras: 'fused' (default True: CUDA-graphed libb200rec step with on-device
sampling), 'sampler' ('device' | 'host': draw triples on device, or iterate the DataLoader like the reference),
'seed', 'eval_precision' (0 exact fp32, 1 tcgen05 bf16 candidates + exact re-score).
SGLTrainer (:432-459) and HALFTrainer (:460-486) ride on the same engine; DOSEaugTrainer (:255-303) and
DOSEdropTrainer (:304-353) drive the DOSE_* models through the library's autograd Functions.
Out of scope: DOSEtest/IDCF/BCE/ML trainers.

What changes underneath: one training step is a replayed CUDA graph (b200rec.engine.BprEngine); evaluation computes
the representation once, then runs the fused score + mask + top-K kernel over large user chunks and a hit-matrix
kernel, instead of a re-propagation, a dense [512, n_items] score matrix, Python exclusion lists and two device
synchronisations per 512 users (trainer.py:150-170).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F
from torch.optim import SGD, Adam  # noqa: F401  (resolved by name from the config, like the reference)
from torch.utils.data import DataLoader, TensorDataset

from b200rec import ops
from b200rec.engine import BprEngine
from dataset import AuxiliaryDataset
from utils import AverageMeter


def get_trainer(config, dataset, model):
    config = config.copy()
    config['dataset'] = dataset
    config['model'] = model
    cls = getattr(sys.modules[__name__], config['name'])
    return cls(config)


class BasicTrainer:
    def __init__(self, trainer_config):
        self.config = trainer_config
        self.name = trainer_config['name']
        self.dataset = trainer_config['dataset']
        self.model = trainer_config['model']
        self.topks = trainer_config['topks']
        self.device = torch.device(trainer_config['device'])
        self.n_epochs = trainer_config['n_epochs']
        self.max_patience = trainer_config.get('max_patience', 50)
        self.val_interval = trainer_config.get('val_interval', 1)
        self.epoch = 0
        self.best_ndcg = -np.inf
        self.save_path = None
        self.opt = None
        # 1 = tcgen05 bf16 candidate pass + exact fp32 re-score (identical ids/scores, D = 64 / 128; falls back to 0 otherwise)
        self.eval_precision = trainer_config.get('eval_precision', 1)
        # users per fused score/top-K call: whole waves of the 128-user CTAs of the tcgen05 kernel (one CTA per SM), so a
        # 2M-user sweep does not run 128 CTAs on 148 SMs chunk after chunk
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == 'cuda' else 128
        self.eval_chunk = max(int(trainer_config['test_batch_size']), 2 * 128 * sms)
        test_user = TensorDataset(torch.arange(self.dataset.n_users, dtype=torch.int64, device=self.device))
        self.test_user_loader = DataLoader(test_user, batch_size=trainer_config['test_batch_size'])
        self.engine = None

    def initialize_optimizer(self):
        opt = getattr(sys.modules[__name__], self.config['optimizer'])
        self.opt = opt(self.model.parameters(), lr=self.config['lr'])

    def train_one_epoch(self):
        raise NotImplementedError

    def record(self, writer, stage, metrics):
        for metric in metrics:
            for k in self.topks:
                writer.add_scalar('{:s}_{:s}/{:s}_{:s}@{:d}'.format(self.model.name, self.name, stage, metric, k),
                                  metrics[metric][k], self.epoch)

    def train(self, verbose=True, writer=None):
        if not self.model.trainable:
            results, metrics, _ = self.eval('val')
            if verbose:
                print('Validation result. {:s}'.format(results))
            return metrics['NDCG'][self.topks[4]]
        os.makedirs('checkpoints', exist_ok=True)
        patience = self.max_patience
        for self.epoch in range(self.n_epochs):
            start_time = time.time()
            self.model.train()
            loss = self.train_one_epoch()
            _, metrics, _ = self.eval('train')
            consumed_time = time.time() - start_time
            if verbose:
                print('Epoch {:d}/{:d}, Loss: {:.6f}, Time: {:.3f}s'.format(self.epoch, self.n_epochs, loss, consumed_time))
            if writer:
                writer.add_scalar('{:s}_{:s}/train_loss'.format(self.model.name, self.name), loss, self.epoch)
                self.record(writer, 'train', metrics)
            if (self.epoch + 1) % self.val_interval != 0:
                continue
            start_time = time.time()
            results, metrics, _ = self.eval('val')
            consumed_time = time.time() - start_time
            if verbose:
                print('Validation result. {:s}Time: {:.3f}s'.format(results, consumed_time))
            if writer:
                self.record(writer, 'validation', metrics)
            ndcg = metrics['NDCG'][self.topks[4]]
            if ndcg > self.best_ndcg:
                if self.save_path:
                    os.remove(self.save_path)
                self.save_path = os.path.join('checkpoints', '{:s}_{:s}_{:s}_{:.3f}.pth'.format(
                    self.model.name, self.name, self.dataset.name, ndcg * 100))
                self.best_ndcg = ndcg
                self.model.save(self.save_path)
                patience = self.max_patience
                print('Best NDCG, save model to {:s}'.format(self.save_path))
            else:
                patience -= self.val_interval
                if patience <= 0:
                    print('Early stopping!')
                    break
        self.model.load(self.save_path)
        return self.best_ndcg

    # ------------------------------------------------------------------------------------------------ metrics
    def _eval_csr(self, eval_data):
        """eval lists as a sorted int32 CSR on the device: (ptr, idx)"""
        dev = self.device
        for split in ('train', 'val', 'test'):  # the dataset's own lists: use its cached device CSR
            if eval_data is getattr(self.dataset, split + '_data', None):
                ptr_d, idx_d = self.dataset.csr(split, device=dev)
                break
        else:
            if hasattr(eval_data, 'ptr'):  # lazy CSR-backed lists
                ptr, idx = np.asarray(eval_data.ptr, dtype=np.int64), np.asarray(eval_data.idx, dtype=np.int64)
            else:
                lens = np.fromiter((len(x) for x in eval_data), dtype=np.int64, count=len(eval_data))
                ptr = np.zeros(len(eval_data) + 1, dtype=np.int64)
                np.cumsum(lens, out=ptr[1:])
                idx = np.concatenate([np.asarray(x, dtype=np.int64) for x in eval_data if len(x)]) if ptr[-1] else \
                    np.zeros(0, dtype=np.int64)
            rows = np.repeat(np.arange(len(ptr) - 1, dtype=np.int64), np.diff(ptr))
            idx = idx[np.lexsort((idx, rows))]
            ptr_d = torch.from_numpy(ptr.astype(np.int32)).to(dev)
            idx_d = torch.from_numpy(idx.astype(np.int32)).to(dev)
        if idx_d.numel() == 0:
            idx_d = torch.zeros(1, dtype=torch.int32, device=dev)
        return ptr_d, idx_d

    def _rec_on_device(self, rec_items):
        rec_d = rec_items if isinstance(rec_items, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rec_items))
        return rec_d.to(device=self.device, dtype=torch.int32).contiguous()

    def hit_matrix(self, eval_data, rec_items):
        """hit[u, j] = rec_items[u, j] in eval_data[u]  (device kernel; eval rows as sorted CSR) and the row lengths"""
        ptr_d, idx_d = self._eval_csr(eval_data)
        n_eval = (ptr_d[1:] - ptr_d[:-1]).cpu().numpy().astype(np.int32)
        return ops.hit_matrix(self._rec_on_device(rec_items), 0, ptr_d, idx_d).cpu().numpy(), n_eval

    def calculate_metrics(self, eval_data, rec_items):
        """Precision / Recall / NDCG @k for every k in topks, averaged over users with at least one eval item
        (trainer.py:115-144) -- one device pass over the recommended ids per 32 cut-offs, 3*len(topks)+1 doubles come back."""
        rec_d = self._rec_on_device(rec_items)
        ks = sorted(set(int(k) for k in self.topks))
        if rec_d.dim() != 2 or ks[0] < 1 or ks[-1] > rec_d.shape[1]:
            raise ValueError('calculate_metrics needs rec_items [n_users, K] with K >= max(topks)')
        results = {'Precision': {}, 'Recall': {}, 'NDCG': {}}
        if rec_d.shape[0] == 0:
            for name in results:
                results[name] = {k: np.float32(np.nan) for k in ks}
            return results
        ptr_d, idx_d = self._eval_csr(eval_data)
        for c0 in range(0, len(ks), 32):  # the kernel takes up to 32 cut-offs per pass
            part = ks[c0:c0 + 32]
            sums = ops.rank_metrics(rec_d, 0, ptr_d, idx_d, part).cpu().numpy()
            n = sums[-1]
            for i, k in enumerate(part):
                for j, name in enumerate(('Precision', 'Recall', 'NDCG')):
                    results[name][k] = np.float32(sums[3 * i + j] / n) if n > 0 else np.float32(np.nan)
        return results

    def recommend_all(self, val_or_test, banned_items=None):
        """top-max(topks) item ids for every user: int32 [n_users, K] on device"""
        self.model.eval()
        dev = self.device
        excl_a = excl_b = None
        if val_or_test != 'train':
            excl_a = self.dataset.csr('train', device=dev)
            if val_or_test == 'test':
                excl_b = self.dataset.csr('val', device=dev)
        banned, kept = None, None
        if banned_items is not None:
            b = np.unique(np.asarray(banned_items).astype(np.int64).ravel())
            if b.size:
                lo, hi = int(b[0]), int(b[-1]) + 1
                if hi - lo == b.size:
                    banned = (lo, hi)  # a contiguous id range (what inductive_eval passes): masked inside the kernel
                else:
                    kept = self._without_items(b, excl_a, excl_b)  # arbitrary index array (trainer.py:166-167)
        k = max(self.topks)
        n_users = self.dataset.n_users
        shard = getattr(self.model, '_dim_shard', None)
        u_lo, u_hi = (0, n_users) if shard is None else shard.user_range(n_users)
        out = []
        with torch.no_grad():
            table_u, table_i = self.model.score_tables()  # collective (column all-gather) when sharded: every rank takes part
            # The fused kernel only emits items that beat a running threshold, so a NaN / -inf score is never returned and
            # such rows come back padded with id -1 (torch.topk would rank NaN first, trainer.py:169).  A diverged model
            # must not pass as "all metrics zero": fail loudly instead.
            if not bool(torch.isfinite(table_u).all()) or not bool(torch.isfinite(table_i).all()):
                raise FloatingPointError('non-finite values in the representation: the model has diverged')
            if kept is not None:
                kept_ids, excl_a, excl_b = kept
                table_i = table_i[kept_ids].contiguous()
            for s in range(u_lo, u_hi, self.eval_chunk):  # multi-GPU: each rank scores its own block of users
                users = torch.arange(s, min(u_hi, s + self.eval_chunk), dtype=torch.int64, device=dev)
                if kept is None:
                    ids, _ = self.model.recommend(users, k, excl_a, excl_b, banned, precision=self.eval_precision)
                else:
                    kk = min(k, int(kept_ids.numel()))
                    ids_c, _ = ops.score_topk(table_u, users, table_i, kk, excl_a, excl_b, None, self.eval_precision)
                    ids = torch.full((users.numel(), k), -1, dtype=torch.int32, device=dev)
                    ids[:, :kk] = torch.where(ids_c >= 0, kept_ids[ids_c.clamp(min=0).long()].to(torch.int32), ids_c)
                out.append(ids)
        ids = torch.cat(out, dim=0) if out else torch.zeros((0, k), dtype=torch.int32, device=dev)
        if shard is not None:
            ids = shard.gather_user_rows(ids, n_users)
        return ids

    def _without_items(self, banned_sorted, excl_a, excl_b):
        """Arbitrary banned ids: score against the item table with those rows REMOVED and map the ids back.  The compact
        ids ascend with the original ones, so the (score desc, id asc) order is unchanged; the exclusion CSRs are
        rewritten into compact ids (banned entries dropped).  Returns (kept original ids int64, excl_a', excl_b')."""
        dev = self.device
        n_items = self.dataset.n_items
        keep = torch.ones(n_items, dtype=torch.bool, device=dev)
        keep[torch.from_numpy(banned_sorted).to(dev)] = False
        kept_ids = torch.nonzero(keep).flatten()
        new_id = (torch.cumsum(keep, 0) - 1).to(torch.int32)

        def remap(excl):
            if excl is None:
                return None
            ptr, idx = excl
            n_u = ptr.numel() - 1
            rows = torch.repeat_interleave(torch.arange(n_u, device=dev), (ptr[1:] - ptr[:-1]).long())
            idx = idx[: rows.numel()].long()
            m = keep[idx]
            counts = torch.bincount(rows[m], minlength=n_u)
            ptr2 = torch.zeros(n_u + 1, dtype=torch.int64, device=dev)
            torch.cumsum(counts, 0, out=ptr2[1:])
            idx2 = new_id[idx[m]].contiguous()
            if idx2.numel() == 0:
                idx2 = torch.zeros(1, dtype=torch.int32, device=dev)
            return ptr2.to(torch.int32), idx2

        return kept_ids, remap(excl_a), remap(excl_b)

    def eval(self, val_or_test, banned_items=None):
        eval_data = getattr(self.dataset, val_or_test + '_data')
        rec_items = self.recommend_all(val_or_test, banned_items)
        metrics = self.calculate_metrics(eval_data, rec_items)
        precison = recall = ndcg = ''
        for k in self.topks:
            precison += '{:.3f}, '.format(metrics['Precision'][k] * 100.)
            recall += '{:.3f}, '.format(metrics['Recall'][k] * 100.)
            ndcg += '{:.3f}, '.format(metrics['NDCG'][k] * 100.)
        results = 'Precision: {:s}Recall: {:s}NDCG: {:s}'.format(precison, recall, ndcg)
        return results, metrics, []

    def inductive_eval(self, n_old_users, n_old_items):
        """Six test passes over {all, old, new} users x {all, old, new} items; ids >= n_old_* are the new nodes."""
        ds = self.dataset
        full = [list(x) for x in ds.test_data]
        n_users, n_items = ds.n_users, ds.n_items

        def run(tag, user_range, item_filter, banned):
            data = [[] for _ in range(n_users)]
            for u in user_range:
                items = full[u]
                if item_filter is not None:
                    items = [i for i in items if item_filter(i)]
                data[u] = items
            ds.test_data = data
            results, metrics, _ = self.eval('test', banned_items=banned)
            print('{:s} result. {:s}'.format(tag, results))
            return metrics

        out = {}
        try:
            out['all_all'] = run('All users and all items', range(n_users), None, None)
            out['old_all'] = run('Old users and all items', range(n_old_users), None, None)
            out['new_all'] = run('New users and all items', range(n_old_users, n_users), None, None)
            out['all_old'] = run('All users and old items', range(n_users), lambda i: i < n_old_items,
                                 np.arange(n_old_items, n_items))
            out['all_new'] = run('All users and new items', range(n_users), lambda i: i >= n_old_items,
                                 np.arange(n_old_items))
            out['old_old'] = run('Old users and old items', range(n_old_users), lambda i: i < n_old_items,
                                 np.arange(n_old_items, n_items))
        finally:
            ds.test_data = full
        return out


class BPRTrainer(BasicTrainer):
    def __init__(self, trainer_config):
        super().__init__(trainer_config)
        self.batch_size = trainer_config['batch_size']
        self.dataloader = DataLoader(self.dataset, batch_size=self.batch_size,
                                     num_workers=trainer_config['dataloader_num_workers'])
        self.initialize_optimizer()
        self.l2_reg = trainer_config['l2_reg']
        self.fused = trainer_config.get('fused', True) and self.config['optimizer'] == 'Adam'
        self.sampler = trainer_config.get('sampler', 'device')
        self.seed = trainer_config.get('seed', 2021)
        self.aux_reg = 0.0

    def _aux_dataset(self):
        return None

    def _engine(self):
        if self.engine is None:
            self.engine = BprEngine(self.model, self.dataset, self.opt, self.batch_size, self.l2_reg,
                                    aux_reg=self.aux_reg, aux_dataset=self._aux_dataset(), seed=self.seed,
                                    partition=self.config.get('partition'),
                                    contrastive_reg=getattr(self, 'contrastive_reg', 0.0))
        return self.engine

    def steps_per_epoch(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def _host_batches(self):
        for batch_data in self.dataloader:
            yield batch_data[:, 0, :].to(dtype=torch.int64)

    def train_one_epoch(self):
        if not self.fused:
            return self._train_one_epoch_autograd()
        eng = self._engine()
        eng.reset_meter()
        if self.sampler == 'host':
            for inputs in self._host_batches():
                if inputs.shape[0] != self.batch_size:  # ragged last batch: one eager step through autograd
                    self._eager_step_in_fused_epoch(eng, inputs.to(self.device), None)
                    continue
                eng.step(host_batch=inputs.pin_memory())
        else:
            for _ in range(self.steps_per_epoch()):
                eng.step()
        eng.sync_optimizer_state()
        return eng.meter_avg()

    def _eager_step_in_fused_epoch(self, eng, inputs, aux_inputs):
        """A batch the captured step cannot take (the ragged tail of a DataLoader epoch) goes through autograd and
        torch.optim.Adam.  The engine keeps the real step count on the device and Adam keeps its own on the host, so the
        count is handed over before the step (bias correction) and the engine's counter and loss meter are advanced
        after it -- the epoch's loss average includes this batch like the reference's AverageMeter (trainer.py:428)."""
        eng.sync_optimizer_state()
        loss = self._autograd_step(inputs, aux_inputs)
        eng.note_external_step(loss, inputs.shape[0])

    # ---- reference-shaped loop through autograd (any optimiser; also the parity path for bpr_forward) ----
    def _loss(self, inputs, aux_inputs):
        users, pos_items, neg_items = inputs[:, 0].contiguous(), inputs[:, 1].contiguous(), inputs[:, 2].contiguous()
        users_r, pos_items_r, neg_items_r, l2_norm_sq = self.model.bpr_forward(users, pos_items, neg_items)
        pos_scores = torch.sum(users_r * pos_items_r, dim=1)
        neg_scores = torch.sum(users_r * neg_items_r, dim=1)
        return F.softplus(neg_scores - pos_scores).mean() + self.l2_reg * l2_norm_sq.mean()

    def _autograd_step(self, inputs, aux_inputs):
        loss = self._loss(inputs, aux_inputs)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        self.model._rep_cache = None
        return loss.item()

    def _train_one_epoch_autograd(self):
        losses = AverageMeter()
        for inputs in self._host_batches():
            inputs = inputs.to(self.device)
            losses.update(self._autograd_step(inputs, None), inputs.shape[0])
        return losses.avg


class IGCNTrainer(BPRTrainer):
    def __init__(self, trainer_config):
        super().__init__(trainer_config)
        self.aux_dataset = AuxiliaryDataset(self.dataset, self.model.user_map, self.model.item_map)
        self.aux_dataloader = DataLoader(self.aux_dataset, batch_size=self.batch_size,
                                         num_workers=trainer_config['dataloader_num_workers'])
        self.aux_reg = trainer_config['aux_reg']

    def _aux_dataset(self):
        return self.aux_dataset

    def _loss(self, inputs, aux_inputs):
        loss = super()._loss(inputs, aux_inputs)
        m = self.model
        users, pos_items, neg_items = aux_inputs[:, 0].contiguous(), aux_inputs[:, 1].contiguous(), aux_inputs[:, 2].contiguous()
        tu = len(m.user_map)
        users_r = ops.gather_rows(m.embedding.weight, users)
        pos_items_r = ops.gather_rows(m.embedding.weight, pos_items, tu)
        neg_items_r = ops.gather_rows(m.embedding.weight, neg_items, tu)
        pos_scores = torch.sum(users_r * pos_items_r * m.w[None, :], dim=1)
        neg_scores = torch.sum(users_r * neg_items_r * m.w[None, :], dim=1)
        return loss + self.aux_reg * F.softplus(neg_scores - pos_scores).mean()

    def train_one_epoch(self):
        if not self.fused:
            losses = AverageMeter()
            for batch_data, a_batch_data in zip(self.dataloader, self.aux_dataloader):
                inputs = batch_data[:, 0, :].to(device=self.device, dtype=torch.int64)
                aux = a_batch_data[:, 0, :].to(device=self.device, dtype=torch.int64)
                losses.update(self._autograd_step(inputs, aux), inputs.shape[0])
            loss = losses.avg
        else:
            eng = self._engine()
            eng.reset_meter()
            eng.refresh_row_scale()
            if self.sampler == 'host':
                for batch_data, a_batch_data in zip(self.dataloader, self.aux_dataloader):
                    inputs = batch_data[:, 0, :].to(dtype=torch.int64)
                    aux = a_batch_data[:, 0, :].to(dtype=torch.int64)
                    if inputs.shape[0] != self.batch_size:
                        self._eager_step_in_fused_epoch(eng, inputs.to(self.device), aux.to(self.device))
                        continue
                    eng.step(host_batch=inputs.pin_memory(), host_aux_batch=aux.pin_memory())
            else:
                for _ in range(self.steps_per_epoch()):
                    eng.step()
            eng.sync_optimizer_state()
            loss = eng.meter_avg()
        self.model.feat_mat_anneal()
        return loss


class SGLTrainer(BPRTrainer):
    """BPR + contrastive_reg * InfoNCE between two edge-dropped views; the views are redrawn after every epoch
    (reference trainer.py:432-459)."""

    def __init__(self, trainer_config):
        super().__init__(trainer_config)
        self.contrastive_reg = trainer_config['contrastive_reg']

    def _loss(self, inputs, aux_inputs):
        users, pos_items, neg_items = inputs[:, 0].contiguous(), inputs[:, 1].contiguous(), inputs[:, 2].contiguous()
        users_r, pos_items_r, neg_items_r, l2_norm_sq, contrastive_loss = self.model.bpr_forward(users, pos_items, neg_items)
        pos_scores = torch.sum(users_r * pos_items_r, dim=1)
        neg_scores = torch.sum(users_r * neg_items_r, dim=1)
        return (F.softplus(neg_scores - pos_scores).mean() + self.l2_reg * l2_norm_sq.mean()
                + self.contrastive_reg * contrastive_loss.mean())

    def train_one_epoch(self):
        loss = super().train_one_epoch()
        self.model.update_aug_adj()
        return loss


class HALFTrainer(SGLTrainer):
    """one augmented view against the full graph (reference trainer.py:460-486)"""


class DOSEaugTrainer(IGCNTrainer):
    """IGCNTrainer + contrastive_reg * InfoNCE between the model's two views (reference trainer.py:255-303): the loop of
    trainer.py:269-293, then feat_mat_anneal() and update_aug_adj() (re-mine, rebuild the edited graph) at the epoch end.
    Runs through the library's autograd Functions (propagation, inductive layer, row gathers, InfoNCE are libb200rec
    kernels with hand-written backwards); the CUDA-graphed engine does not carry this model's two dropout draws per step."""

    def __init__(self, trainer_config):
        super().__init__(trainer_config)
        self.contrastive_reg = trainer_config['contrastive_reg']
        self.fused = False

    def _loss(self, inputs, aux_inputs):
        users, pos_items, neg_items = inputs[:, 0].contiguous(), inputs[:, 1].contiguous(), inputs[:, 2].contiguous()
        users_r, pos_items_r, neg_items_r, l2_norm_sq, contrastive_loss = self.model.bpr_forward(users, pos_items, neg_items)
        pos_scores = torch.sum(users_r * pos_items_r, dim=1)
        neg_scores = torch.sum(users_r * neg_items_r, dim=1)
        bpr_loss = F.softplus(neg_scores - pos_scores).mean()
        m = self.model
        a_users, a_pos, a_neg = aux_inputs[:, 0].contiguous(), aux_inputs[:, 1].contiguous(), aux_inputs[:, 2].contiguous()
        tu = len(m.user_map)
        e_u = ops.gather_rows(m.embedding.weight, a_users)
        e_p = ops.gather_rows(m.embedding.weight, a_pos, tu)
        e_n = ops.gather_rows(m.embedding.weight, a_neg, tu)
        aux_loss = F.softplus(torch.sum(e_u * e_n * m.w[None, :], dim=1) - torch.sum(e_u * e_p * m.w[None, :], dim=1)).mean()
        return bpr_loss + self.l2_reg * l2_norm_sq.mean() + self.aux_reg * aux_loss + self.contrastive_reg * contrastive_loss.mean()

    def train_one_epoch(self):
        loss = super().train_one_epoch()  # the autograd loop + feat_mat_anneal()
        self.model.update_aug_adj()
        return loss


class DOSEdropTrainer(DOSEaugTrainer):
    """reference trainer.py:304-353 (the same loop; the model's update_aug_adj rebuilds the dropped graph)"""

