"""Dataset ingest: raw D4RL-style arrays -> transitions -> device-resident replay buffers.

Mirror of dataset/call_dataset.py:21-110 (``call_tar_dataset``): the reference reads ``observations, actions, rewards,
terminals[, timeouts]`` from an HDF5 file and walks the N - 1 first rows in a Python loop, pairing row i with row i + 1 as
its next observation (also across episode boundaries -- the loop only *counts* episodes, :82-91) and flattening [N, 1]
rewards.  Here the walk is five array slices; ``load_buffers`` hands the result to ReplayBuffer.convert_D4RL, i.e. one
H2D copy + one pack_rows launch per buffer.  h5py / gym / d4rl are not part of this image: the HDF5 reader is used when
h5py is importable and says so when it is not; ``.npz`` files with the same keys are read natively.
"""
import os

import numpy as np

RAW_KEYS = ("observations", "actions", "rewards", "terminals")


def read_raw(path):
    """Every dataset of an .hdf5 file (call_dataset.py:11-19, 51-57) or every array of an .npz file, as a dict."""
    if path.endswith(".npz"):
        with np.load(path) as f:
            return {k: f[k] for k in f.files}
    try:
        import h5py
    except ImportError as e:           # no CPU/alternative reader is faked: say what is missing
        raise ImportError("mobody_b200.dataset.read_raw: reading .hdf5 needs h5py (not installed in this image); "
                          "convert the file to .npz with the same keys or install h5py") from e
    out = {}
    with h5py.File(path, "r") as f:
        def visitor(name, item):
            if isinstance(item, h5py.Dataset):
                try:
                    out[name] = item[:]
                except ValueError:
                    out[name] = item[()]
        f.visititems(visitor)
    return out


def transitions_from_raw(dataset):
    """call_dataset.py:59-110 without the Python loop: rows 0 .. N-2, next_observations = observations shifted by one row,
    rewards flattened, terminals as bool.  Returns the reference's dict (float32 arrays, ``terminals`` bool)."""
    for k in RAW_KEYS:
        if k not in dataset:
            raise KeyError(f"dataset has no '{k}'")
    n = dataset["rewards"].shape[0]
    obs = np.asarray(dataset["observations"])
    rew = np.asarray(dataset["rewards"])
    if rew.ndim > 1:
        rew = rew.reshape(rew.shape[0], -1)[:, 0]                      # `.astype(np.float32)[0]` of a [1] row (:73-76)
    m = max(n - 1, 0)
    return {
        "observations": obs[:m].astype(np.float32),
        "actions": np.asarray(dataset["actions"])[:m].astype(np.float32),
        "next_observations": obs[1:m + 1].astype(np.float32),
        "rewards": rew[:m].astype(np.float32),
        "terminals": np.asarray(dataset["terminals"])[:m].astype(bool),
    }


def tar_dataset_path(root, tar_env_name, shift_scale, quality="random"):
    """File the reference opens for a target-domain dataset (call_dataset.py:21-49)."""
    name = tar_env_name.replace("-", "_")
    if any(k in name for k in ("halfcheetah", "hopper", "walker2d")) or name.split("_")[0] == "ant":
        return os.path.join(root, "mujoco", f"{name}_{shift_scale}_{quality}.hdf5")
    if any(k in name for k in ("pen", "door", "relocate", "hammer")):
        return os.path.join(root, "adroit", f"{name}_{shift_scale}_{quality}.hdf5")
    if "antmaze" in name:
        return os.path.join(root, "antmaze", f"{name}_{shift_scale}.hdf5")
    raise NotImplementedError(tar_env_name)


def call_tar_dataset(tar_env_name, shift_scale, quality="random", root=None):
    """Reference entry point (call_dataset.py:21): the target-domain transitions of ``<root>/<domain>/<name>_<shift>[_<quality>].hdf5``
    (or the .npz next to it)."""
    root = root or os.path.join(os.getcwd(), "dataset")
    path = tar_dataset_path(root, tar_env_name, shift_scale, quality)
    if not os.path.exists(path) and os.path.exists(path[:-5] + ".npz"):
        path = path[:-5] + ".npz"
    return transitions_from_raw(read_raw(path))


def load_buffers(src_dataset, tar_dataset, state_dim, action_dim, device, seed=0):
    """(src_replay_buffer, tar_replay_buffer) resident on ``device`` (train_mobody.py:640-660: two ReplayBuffers filled by
    convert_D4RL)."""
    from .buffer import ReplayBuffer
    src, tar = ReplayBuffer(state_dim, action_dim, device, seed=seed), ReplayBuffer(state_dim, action_dim, device, seed=seed + 1)
    src.convert_D4RL(src_dataset)
    tar.convert_D4RL(tar_dataset)
    return src, tar
