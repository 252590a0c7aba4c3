"""Drop-in for the reference's config.py: per-dataset lists of (dataset_config, model_config, trainer_config) triples
with the reference's keys and shipped hyper-parameters for the in-scope models (config.py:1-23 Gowalla; the Yelp and
Amazon blocks :103-125, :206-228 use the same model/trainer values).  Entries for out-of-scope models (ItemKNN, NGCF,
DOSE_*, MultiVAE, NeuMF ...) are not reproduced.  `get_synthetic_config` serves the BASELINE.json shapes.
"""
TOPKS = [1, 5, 10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100]


def _triples(dataset_config, device):
    out = []
    common = {'device': device, 'n_epochs': 1000, 'batch_size': 2048, 'dataloader_num_workers': 6,
              'test_batch_size': 512, 'topks': TOPKS}
    out.append((dataset_config,
                {'name': 'MF', 'embedding_size': 64, 'device': device},
                dict(common, name='BPRTrainer', optimizer='Adam', lr=1.e-4, l2_reg=1.e-3)))
    out.append((dataset_config,
                {'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 3, 'device': device},
                dict(common, name='BPRTrainer', optimizer='Adam', lr=1.e-3, l2_reg=1.e-4)))
    out.append((dataset_config,
                {'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': device, 'dropout': 0.3,
                 'feature_ratio': 1},
                dict(common, name='IGCNTrainer', optimizer='Adam', lr=1.e-3, l2_reg=0., aux_reg=0.01)))
    out.append((dataset_config,
                {'name': 'IMF', 'embedding_size': 64, 'n_layers': 0, 'device': device, 'dropout': 0.3,
                 'feature_ratio': 1},
                dict(common, name='IGCNTrainer', optimizer='Adam', lr=1.e-3, l2_reg=0., aux_reg=0.01)))
    return out


def get_gowalla_config(device):
    return _triples({'name': 'ProcessedDataset', 'path': 'data/Gowalla/time', 'device': device}, device)


def get_yelp_config(device):
    return _triples({'name': 'ProcessedDataset', 'path': 'data/Yelp/time', 'device': device}, device)


def get_amazon_config(device):
    return _triples({'name': 'ProcessedDataset', 'path': 'data/Amazon/time', 'device': device}, device)


def get_synthetic_config(device, shape='c2', seed=0):
    """BASELINE.json shapes ('c1' Gowalla-, 'c2' Yelp2018-, 'c3' Amazon-book-shaped, 'c4' 2M x 1M power-law)."""
    from b200rec.synth import SHAPES
    _, _, _, d, n_layers = SHAPES[shape]
    triples = _triples({'name': 'SyntheticDataset', 'shape': shape, 'seed': seed, 'device': device}, device)
    for _, model_cfg, _ in triples:
        model_cfg['embedding_size'] = d
        if model_cfg['name'] in ('LightGCN', 'IGCN'):
            model_cfg['n_layers'] = n_layers
    return triples
