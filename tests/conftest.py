import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['meta']))
    return meta, {k: z[k] for k in z.files if k != 'meta'}


def weight_fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


GOLDEN_CASES = ['tasnet_r2', 'spe_cat_r2_eval', 'spe_add_r2_eval', 'spe_mul_r2_eval', 'spe_film_r2_eval',
                'spe_att_r2_eval', 'spe_cat_r2_train', 'spe_att_r2_train_b1', 'spe_film_gln_relu_r2',
                'spe_cat_uni_r2', 'ira_cat_r2_eval', 'ira_cat_r2_train', 'spe_cat_r6_3s', 'tasnet_r6_3s',
                'speech_att_r6', 'speech_cat_r6_wx3']       # the last two: real speech (example.ipynb cell 15)
