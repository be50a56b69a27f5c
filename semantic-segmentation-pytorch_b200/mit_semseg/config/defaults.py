"""The global configuration tree `cfg` with the reference's keys and default values (mit_semseg/config/defaults.py:7-97):
DIR, DATASET, MODEL, TRAIN, VAL, TEST. yacs' CfgNode when it is installed, the bundled work-alike otherwise."""
try:
    from yacs.config import CfgNode as CN
except ImportError:   # not part of this image
    from .cfgnode import CfgNode as CN

_DEFAULTS = {
    "DIR": "ckpt/ade20k-resnet50dilated-ppm_deepsup",
    "DATASET": {
        "root_dataset": "./data/", "list_train": "./data/training.odgt", "list_val": "./data/validation.odgt",
        "num_class": 150,
        "imgSizes": (300, 375, 450, 525, 600),   # short-edge sizes drawn per batch (train) / evaluated one by one (val, test)
        "imgMaxSize": 1000,                      # cap on the long edge
        "padding_constant": 8,                   # batches are padded to a multiple of the network's largest stride
        "segm_downsampling_rate": 8,             # label maps are produced at 1/8 resolution
        "random_flip": True,
    },
    "MODEL": {"arch_encoder": "resnet50dilated", "arch_decoder": "ppm_deepsup", "weights_encoder": "", "weights_decoder": "",
              "fc_dim": 2048},
    "TRAIN": {
        "batch_size_per_gpu": 2, "num_epoch": 20, "start_epoch": 0, "epoch_iters": 5000,
        "optim": "SGD", "lr_encoder": 0.02, "lr_decoder": 0.02, "lr_pow": 0.9, "beta1": 0.9, "weight_decay": 1e-4,
        "deep_sup_scale": 0.4, "fix_bn": False, "workers": 16, "disp_iter": 20, "seed": 304,
    },
    "VAL": {"batch_size": 1, "visualize": False, "checkpoint": "epoch_20.pth"},
    "TEST": {"batch_size": 1, "checkpoint": "epoch_20.pth", "result": "./"},
}


def _tree(d):
    node = CN()
    for k, v in d.items():
        setattr(node, k, _tree(v) if isinstance(v, dict) else v)
    return node


_C = _tree(_DEFAULTS)
