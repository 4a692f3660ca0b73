"""On-disk frames of the reference's datasets and roll-outs: one `velocity_%06d.npz` /
`pressure_%06d.npz` pair per frame, each holding `arr_0` = staggered tensor [1, ny+1, nx+1, 2] / centred [1, ny, nx, 1]
(spatial_mixing_layer.py:60-75), and the training-sample assembly of diffpiso/datamanagement.py:25-64."""
import os
from collections.abc import Iterable

import numpy as np
import torch


def create_base_dir(path, name):
    """datamanagement.py:11-22: first free `path + name + %06d` directory."""
    i = 0
    while os.path.exists(path + name + str(i).zfill(6)):
        i += 1
    os.makedirs(path + name + str(i).zfill(6))
    return path + name + str(i).zfill(6)


def frame_path(directory, field_name, frame):
    return os.path.join(directory, "%s_%s.npz" % (field_name, str(int(frame)).zfill(6)))


def save_frame(directory, frame, velocity, pressure):
    """Write one frame the way the reference's roll-outs do (np.savez -> `arr_0`); batch must be 1 per file."""
    from .grids import CenteredGrid, StaggeredGrid
    v = velocity.staggered_tensor() if isinstance(velocity, StaggeredGrid) else velocity
    p = pressure.data if isinstance(pressure, CenteredGrid) else pressure
    v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    p = p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p)
    if v.shape[0] != 1 or p.shape[0] != 1:
        raise ValueError("one sample per frame file (arr_0 is [1, ...])")
    np.savez(frame_path(directory, "velocity", frame), v)
    np.savez(frame_path(directory, "pressure", frame), p)


def load_frame(directory, frame, field_names=("velocity", "pressure")):
    return tuple(np.load(frame_path(directory, n, frame))["arr_0"].astype(np.float32) for n in field_names)


def data_path_assembler(paths, field_names, characteristics, start_frame, frame_count, step_count, dt_ratio=1):
    """datamanagement.py:35-48: per training sample the file names of frames i, i+dt_ratio, ..., i+step_count*dt_ratio
    of every field, plus the sample's characteristics entry."""
    file_list = tuple([[] for _ in range(len(field_names) + 1)])
    for p, pth in enumerate(paths):
        for i in range(start_frame[p], start_frame[p] + frame_count[p] - step_count[p] * dt_ratio):
            for n, name in enumerate(field_names):
                file_list[n].append([pth + name + "_" + str(i + j * dt_ratio).zfill(6) + ".npz"
                                     for j in range(0, step_count[p] + 1)])
            if isinstance(characteristics[p], Iterable):
                file_list[-1].append(characteristics[p][i - start_frame[p]])
            else:
                file_list[-1].append(characteristics[p])
    return file_list


def load_function(*data_tuple):
    """datamanagement.py:51-58: stack the frames of one sample along a new axis 1 -> [1, steps+1, ...] per field, and
    the characteristics as [1, k] float32."""
    output = []
    for d in range(len(data_tuple) - 1):
        output.append(np.concatenate([np.expand_dims(np.load(f)["arr_0"].astype(np.float32), axis=1)
                                      for f in data_tuple[d]], axis=1))
    output.append(np.expand_dims(np.array(data_tuple[-1]), 0).astype(np.float32))
    return tuple(output)


class FrameDataset(torch.utils.data.Dataset):
    """The sample list of `data_path_assembler` as a torch dataset (replaces make_tf_dataset/load_function_wrapper,
    datamanagement.py:25-32,61-64); `rank`/`world_size` take every world_size-th sample for batch sharding."""

    def __init__(self, file_tuple, rank=0, world_size=1):
        n = len(file_tuple[0])
        self.index = list(range(rank, n, world_size))
        self.files = file_tuple

    def __len__(self):
        return len(self.index)

    def __getitem__(self, k):
        i = self.index[k]
        return tuple(torch.as_tensor(np.asarray(a[0])) for a in load_function(*[f[i] for f in self.files]))
