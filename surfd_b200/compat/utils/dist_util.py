"""utils/dist_util.py of the reference: setup_dist records the device id (a no-op there too, :18-41), dev() returns it."""
import torch as th

used_device = 0


def setup_dist(device=0):
    global used_device
    used_device = device


def dev():
    if th.cuda.is_available() and used_device >= 0:
        return th.device(f"cuda:{used_device}")
    return th.device("cpu")


def load_state_dict(path, **kwargs):
    return th.load(path, **kwargs)
