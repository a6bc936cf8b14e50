"""The two helpers of the reference's deeplens/utils.py that its driver scripts call around the hot path (1_fit_psfnet.py:10-17,
2_dfdp_net.py): `set_seed` (utils.py:136-145) and `set_logger` (utils.py:148-164).  The image-quality metrics and plotting helpers of
that module (lpips, skimage, OpenCV) are outside the hot path."""
import logging
import os
import random

import numpy as np
import torch


def set_seed(seed=0):
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.enabled = False


def set_logger(dir="./"):
    logger = logging.getLogger()
    logger.setLevel("DEBUG")
    formatter = logging.Formatter("%(asctime)s:%(levelname)s:%(message)s", "%Y-%m-%d %H:%M:%S")
    chlr = logging.StreamHandler()
    chlr.setFormatter(formatter)
    chlr.setLevel("INFO")
    fhlr = logging.FileHandler(f"{dir}/output.log")
    fhlr.setFormatter(formatter)
    fhlr.setLevel("INFO")
    logger.addHandler(chlr)
    logger.addHandler(fhlr)
