"""Drop-in for the reference's config.py: per-dataset lists of (dataset_config, model_config, trainer_config) triples
with the reference's keys, shipped hyper-parameters AND list positions (scripts index the lists, e.g.
`get_gowalla_config(device)[2]`; /root/reference/config.py:1-100 Gowalla, :103-203 Yelp, :206-289 Amazon, :292-408
alibaba, :411-... ml-1m).  Every entry of the reference's lists is present -- entries naming models this package does not
build (ItemKNN, NGCF, MultiVAE, IMCGAE, IDCF_LGCN, NeuMF: SURVEY.md section 2.1, out of scope) are plain data and fail at
get_model with the class name.  `tests/golden/config_ref.json` holds the reference's own lists and
`tests/test_utils_cpu.py` compares.  `get_synthetic_config` serves the BASELINE.json shapes.

Stated here as one recipe table plus per-dataset differences rather than five hand-expanded lists.
"""
import copy

TOPKS = [1, 5, 10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100]
_COMMON = {'n_epochs': 1000, 'batch_size': 2048, 'dataloader_num_workers': 6, 'test_batch_size': 512}


def _adam(name, lr, l2_reg, **extra):
    return dict(name=name, optimizer='Adam', lr=lr, l2_reg=l2_reg, **extra)


def _dose(model, trainer, aug_num, **model_extra):
    return (dict(name=model, embedding_size=64, n_layers=3, dropout=0.3, feature_ratio=1, aug_num=aug_num, **model_extra),
            _adam(trainer, 1.e-3, 0., contrastive_reg=1.e-1, aux_reg=0.001))


# (model_config, trainer_config[, dataset extras]) in the reference's Gowalla order; `device`, the common trainer keys
# and `topks` are added by _expand
_RECIPES = [
    (dict(name='MF', embedding_size=64), _adam('BPRTrainer', 1.e-4, 1.e-3)),
    (dict(name='LightGCN', embedding_size=64, n_layers=3), _adam('BPRTrainer', 1.e-3, 1.e-4)),
    (dict(name='IGCN', embedding_size=64, n_layers=3, dropout=0.3, feature_ratio=1),
     _adam('IGCNTrainer', 1.e-3, 0., aux_reg=0.01)),
    (dict(name='ItemKNN', k=1000), dict(name='BasicTrainer', n_epochs=0, _bare=True)),
    (dict(name='NGCF', embedding_size=64, layer_sizes=[64, 64, 64], dropout=0.1), _adam('BPRTrainer', 1.e-3, 1.e-3)),
    (dict(name='MultiVAE', layer_sizes=[64, 32], dropout=0.7), _adam('MLTrainer', 1.e-3, 1.e-4, kl_reg=0.2, batch_size=512)),
    (dict(name='IMF', embedding_size=64, n_layers=0, dropout=0.1, feature_ratio=1.),
     _adam('IGCNTrainer', 1.e-3, 1.e-5, aux_reg=0.1)),
    (dict(name='IMCGAE', embedding_size=64, n_layers=3, dropout=0.3), _adam('BPRTrainer', 1.e-3, 0.)),
    (dict(name='IDCF_LGCN', embedding_size=64, n_layers=3, n_headers=4, lgcn_path='lgcn.pth'),
     _adam('IDCFTrainer', 1.e-3, 1.e-4, contrastive_reg=1.e-3)),
    (dict(name='NeuMF', embedding_size=64, layer_sizes=[64, 64, 64]),
     _adam('BCETrainer', 1.e-3, 1.e-3, test_batch_size=64, mf_pretrain_epochs=100, mlp_pretrain_epochs=100, max_patience=100),
     dict(neg_ratio=4)),   # the reference copies dataset_config here and keeps the copy for every later entry
    _dose('DOSE_aug', 'DOSEaugTrainer', 500000),
    _dose('DOSE_drop3', 'DOSEdropTrainer', 500000, aug_rate=0.5),
    _dose('DOSE_aug_drop2', 'DOSEdropTrainer', 100000),
]


def _expand(path, device, recipes):
    dataset_config = {'name': 'ProcessedDataset', 'path': path, 'device': device}
    out = []
    for rec in recipes:
        model, trainer = copy.deepcopy(rec[0]), copy.deepcopy(rec[1])
        if len(rec) > 2:
            dataset_config = dict(dataset_config, **rec[2])
        model['device'] = device
        bare = trainer.pop('_bare', False)
        full = {'device': device, 'topks': list(TOPKS)}
        if not bare:
            full.update(_COMMON)
        else:
            full['test_batch_size'] = _COMMON['test_batch_size']
        full.update(trainer)
        out.append((dataset_config, model, full))
    return out


def _patched(changes, drop=(), replace=None):
    """the Gowalla recipes with per-index (model_changes, trainer_changes) applied, entries dropped / replaced"""
    recs = [tuple(copy.deepcopy(x) for x in r) for r in _RECIPES]
    for i, (mc, tc) in changes.items():
        recs[i][0].update(mc)
        recs[i][1].update(tc)
    for i, r in (replace or {}).items():
        recs[i] = r
    return [r for i, r in enumerate(recs) if i not in drop]


def get_gowalla_config(device):
    return _expand('data/Gowalla/time', device, _RECIPES)


def get_yelp_config(device):
    recs = _patched({0: ({}, {'lr': 1.e-3}), 4: ({'dropout': 0.3}, {}),
                     9: ({}, {'lr': 1.e-2, 'l2_reg': 1.e-2, 'topks': [20]}),
                     10: ({'aug_num': 800000}, {}), 11: ({'aug_num': 1000000, 'aug_rate': 0.7}, {}),
                     12: ({'aug_num': 300000}, {})},
                    replace={6: (dict(name='DOSE_drop2', embedding_size=64, n_layers=3, dropout=0.3, feature_ratio=1,
                                      aug_num=500000, aug_rate=0.5), _adam('IGCNTrainer', 1.e-3, 1.e-5, aux_reg=0.01))})
    return _expand('data/Yelp/time', device, recs)


def get_amazon_config(device):
    recs = _patched({0: ({}, {'lr': 1.e-3, 'l2_reg': 1.e-4}), 1: ({}, {'l2_reg': 1.e-5}), 2: ({'dropout': 0.0}, {}),
                     3: ({'k': 10}, {}), 4: ({'dropout': 0.3}, {'l2_reg': 1.e-4}), 5: ({}, {'l2_reg': 1.e-5}),
                     6: ({'dropout': 0.3}, {}), 7: ({'dropout': 0.9}, {}), 10: ({'aug_num': 1000000}, {}),
                     12: ({'aug_num': 1000000}, {})},
                    drop=(8, 9),
                    replace={11: (dict(name='DOSE_aug', embedding_size=64, n_layers=3, dropout=0.3, feature_ratio=0.6,
                                       aug_num=1000000, aug_rate=0.7),
                                  _adam('DOSEdropTrainer', 1.e-3, 0., contrastive_reg=1.e-1, aux_reg=0.001))})
    return _expand('data/Amazon/time', device, recs)


def get_alibaba_config(device):
    return _expand('data/alibaba/time', device, _RECIPES)


def get_ml_config(device):
    return _expand('data/ml-1m/time', device, _RECIPES)


def get_synthetic_config(device, shape='c2', seed=0):
    """BASELINE.json shapes ('c1' Gowalla-, 'c2' Yelp2018-, 'c3' Amazon-book-shaped, 'c4' 2M x 1M power-law): the in-scope
    entries (MF, LightGCN, IGCN, IMF) on a SyntheticDataset"""
    from b200rec.synth import SHAPES
    _, _, _, d, n_layers = SHAPES[shape]
    keep = [r for r in _RECIPES if r[0]['name'] in ('MF', 'LightGCN', 'IGCN', 'IMF')]
    triples = _expand('', device, keep)
    ds_cfg = {'name': 'SyntheticDataset', 'shape': shape, 'seed': seed, 'device': device}
    out = []
    for _, model_cfg, trainer_cfg in triples:
        model_cfg['embedding_size'] = d
        if model_cfg['name'] in ('LightGCN', 'IGCN'):
            model_cfg['n_layers'] = n_layers
        out.append((ds_cfg, model_cfg, trainer_cfg))
    return out
