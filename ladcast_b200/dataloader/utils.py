"""Tensor-level pieces of `ladcast.dataloader.utils` that the rollout callers use (reference
dataloader/utils.py:223-306 and the static-field preparation of pipelines/pred_rollout.py:260-291).  The xarray / zarr
front end (`xarr_to_tensor`, `tensor_to_xarr`, `filter_time_range`) is out of scope.

Inside `roll_out_latent` / `encode_fused` / `decode_fused` these transforms are fused into CUDA epilogues; the functions
here are the reference-compatible host API for everything around that."""
from typing import Dict, List, Optional, Tuple

import torch

# channel order of the 84 field channels: 6 atmospheric variables x 13 levels, then 6 surface variables
# (pipelines/pred_rollout.py:33-46); sea_surface_temperature is channel 82
VAR_LIST = [
    "geopotential", "specific_humidity", "temperature", "u_component_of_wind", "v_component_of_wind",
    "vertical_velocity", "10m_u_component_of_wind", "10m_v_component_of_wind", "2m_temperature",
    "mean_sea_level_pressure", "sea_surface_temperature", "total_precipitation_6hr",
]


def _as_tensors(sample: torch.Tensor, mean, std) -> Tuple[torch.Tensor, torch.Tensor]:
    if not isinstance(mean, torch.Tensor):
        mean = torch.tensor(mean, device=sample.device)
        std = torch.tensor(std, device=sample.device)
    return mean, std


def normalize_transform_3D(sample: torch.Tensor, mean, std, target_std=1):
    """(C, T, H, W): (x - mean_c) / std_c * target_std   (dataloader/utils.py:223-230)."""
    mean, std = _as_tensors(sample, mean, std)
    return ((sample - mean[:, None, None, None]) / std[:, None, None, None]) * target_std


def inverse_normalize_transform_3D(normalized_sample: torch.Tensor, mean, std, target_std=1):
    """(C, T, H, W): x / target_std * std_c + mean_c   (dataloader/utils.py:233-240)."""
    mean, std = _as_tensors(normalized_sample, mean, std)
    return (normalized_sample / target_std) * std[:, None, None, None] + mean[:, None, None, None]


def get_transform_3D(transform: Optional[str], transform_args: Optional[dict]):
    """dataloader/utils.py:243-255."""
    if transform == "normalize":
        mean, std = transform_args["mean"], transform_args["std"]
        if "target_std" in transform_args:
            target_std = transform_args["target_std"]
            return lambda x: normalize_transform_3D(x, mean, std, target_std)
        return lambda x: normalize_transform_3D(x, mean, std)
    if transform is None:
        return lambda x: x
    raise NotImplementedError(f"Transform: {transform} not implemented.")


def get_inv_transform_3D(transform: Optional[str], transform_args: Optional[dict]):
    """dataloader/utils.py:258-269."""
    if transform == "normalize":
        mean, std = transform_args["mean"], transform_args["std"]
        if "target_std" in transform_args:
            target_std = transform_args["target_std"]
            return lambda x: inverse_normalize_transform_3D(x, mean, std, target_std)
        return lambda x: inverse_normalize_transform_3D(x, mean, std)
    if transform is None:
        return lambda x: x
    raise NotImplementedError(f"Transform: {transform} not implemented.")


def precompute_mean_std(normalization_param_dict: Dict, variable_names: List[str]):
    """Per-channel mean / std tensors in `variable_names` order; variables with per-level statistics contribute one
    entry per level in the dict's level order (dataloader/utils.py:272-306)."""
    mean_list, std_list = [], []
    for var_name in variable_names:
        if var_name not in normalization_param_dict:
            raise ValueError(f"No normalization parameters found for variable {var_name}.")
        norm_params = normalization_param_dict[var_name]
        if isinstance(norm_params["mean"], dict):
            for level in norm_params["mean"].keys():
                mean_list.append(norm_params["mean"][level])
                std_list.append(norm_params["std"][level])
        else:
            mean_list.append(norm_params["mean"])
            std_list.append(norm_params["std"])
    return torch.tensor(mean_list), torch.tensor(std_list)


def prepare_static_conditioning(lsm: Optional[torch.Tensor], orography: Optional[torch.Tensor],
                                crop_south_pole: bool = True) -> Optional[torch.Tensor]:
    """Static encoder channels as pipelines/pred_rollout.py:260-291 builds them: land-sea mask (lat, lon) and the four
    orography fields (4, lat, lon) on the 121-row grid, south-pole row dropped, concatenated to (C_s, lat, lon) and
    standardised per channel over (lat, lon) with the unbiased std."""
    parts = []
    if lsm is not None:
        parts.append((lsm[1:, :] if crop_south_pole else lsm).unsqueeze(0))
    if orography is not None:
        parts.append(orography[:, 1:, :] if crop_south_pole else orography)
    if not parts:
        return None
    static = torch.cat(parts, dim=0).float()
    mean = static.mean(dim=(1, 2), keepdim=True)
    std = static.std(dim=(1, 2), keepdim=True)
    return (static - mean) / std
