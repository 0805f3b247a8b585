"""Observation slicing contract of the reference (src/utils.py:37-48)."""
import numpy as np
import torch


def convert_obs_to_dict(config, observation):
    V, P = config.vms, config.pms
    placement = observation[:V].to(int) if isinstance(observation, torch.Tensor) else np.asarray(observation[:V]).astype(int)
    return dict(vm_placement=placement, vm_cpu=observation[V:2 * V], vm_memory=observation[2 * V:3 * V],
                cpu=observation[3 * V:3 * V + P], memory=observation[3 * V + P:])
