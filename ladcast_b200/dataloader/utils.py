"""Tensor-level pieces of `ladcast.dataloader.utils` that the rollout callers use (reference
dataloader/utils.py:223-306 and the static-field preparation of pipelines/pred_rollout.py:260-291).  The xarray / zarr
front end is replaced by an array front end with the same channel layout and NaN conventions: a
"field set" is a plain mapping {variable name -> array}, atmospheric variables (time, level, lat, lon), surface
variables (time, lat, lon) — what `np.load(..., mmap_mode="r")` of per-variable .npy files or a zarr group's arrays
give — see `fields_to_tensor` / `tensor_to_fields` / `select_init_times`.

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


SST_FILL_VALUE = -2.0  # normalised sea-surface temperature over land (dataloader/utils.py:399-404, weather_dataset.py)


def _to_tensor(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a
    import numpy as np

    return torch.from_numpy(np.ascontiguousarray(a))


def fields_to_tensor(fields: Dict, variable_names: Optional[List[str]] = None, level_index: Optional[List[int]] = None,
                     normalization_param_dict: Optional[Dict] = None, mean_tensor: Optional[torch.Tensor] = None,
                     std_tensor: Optional[torch.Tensor] = None, static_conditioning_tensor: Optional[torch.Tensor] = None,
                     crop_south_pole: bool = False) -> torch.Tensor:
    """Array counterpart of `xarr_to_tensor` (reference dataloader/utils.py:357-448): stacks a field set into the
    model's (C, T, H, W) layout — variables in `variable_names` order (default: VAR_LIST entries present), an
    atmospheric variable contributing one channel per selected level, time-independent entries skipped — then
    standardises per channel, replaces the NaNs of `sea_surface_temperature` (land) by -2 AFTER normalisation and
    optionally appends the static conditioning channels (C_s, H, W) broadcast over time.
    level_index: positions along the level axis to keep (the reference selects by level value on the xarray side).
    crop_south_pole: drop latitude row 0 of a 121-row grid (pred_rollout.py crops `[1:]`)."""
    names = [n for n in (variable_names or VAR_LIST) if n in fields]
    parts, sst_channel, c = [], None, 0
    for name in names:
        a = _to_tensor(fields[name]).to(torch.float32)
        if a.dim() == 4:  # (time, level, lat, lon) -> (level, time, lat, lon)
            if level_index is not None:
                a = a[:, level_index]
            a = a.permute(1, 0, 2, 3)
        elif a.dim() == 3:  # (time, lat, lon)
            if name == "sea_surface_temperature":
                sst_channel = c
            a = a.unsqueeze(0)
        else:
            continue  # static (lat, lon) entries such as land_sea_mask have no time axis
        parts.append(a)
        c += a.shape[0]
    if not parts:
        raise ValueError("no time-dependent variable found in the field set")
    x = torch.cat(parts, dim=0)
    if crop_south_pole:
        x = x[:, :, 1:]
    if normalization_param_dict is not None:
        mean_tensor, std_tensor = precompute_mean_std(normalization_param_dict,
                                                      [n for n in names if n != "land_sea_mask" and _to_tensor(fields[n]).dim() >= 3])
    if mean_tensor is not None:
        if mean_tensor.numel() != x.shape[0]:
            raise ValueError(f"{mean_tensor.numel()} channel statistics for {x.shape[0]} channels")
        x = normalize_transform_3D(x, mean_tensor.to(x.dtype), std_tensor.to(x.dtype))
        if sst_channel is not None:
            x[sst_channel] = torch.nan_to_num(x[sst_channel], nan=SST_FILL_VALUE)
    if static_conditioning_tensor is not None:
        st = static_conditioning_tensor.to(x.dtype).unsqueeze(1).expand(-1, x.shape[1], -1, -1)
        x = torch.cat([x, st], dim=0)
    return x


def tensor_to_fields(x: torch.Tensor, variable_names: List[str], n_levels: Dict[str, int],
                     normalization_param_dict: Optional[Dict] = None, mean_tensor: Optional[torch.Tensor] = None,
                     std_tensor: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Array counterpart of `tensor_to_xarr` (reference dataloader/utils.py:452-513): (C, T, H, W) -> field set on the
    host, de-normalised first if statistics are given.  n_levels[name] = number of level channels of an atmospheric
    variable (absent / 0 = surface variable)."""
    if normalization_param_dict is not None:
        mean_tensor, std_tensor = precompute_mean_std(normalization_param_dict, variable_names)
    if mean_tensor is not None:
        x = inverse_normalize_transform_3D(x, mean_tensor.to(x.device), std_tensor.to(x.device))
    x = x.cpu()
    out, c = {}, 0
    for name in variable_names:
        k = int(n_levels.get(name, 0))
        if k > 0:
            out[name] = x[c : c + k].permute(1, 0, 2, 3)  # (time, level, lat, lon)
            c += k
        else:
            out[name] = x[c]
            c += 1
    if c != x.shape[0]:
        raise ValueError("Mismatch in the number of variables.")
    return out


def select_init_times(times, num_samples_per_month: int, enforce_year=None) -> List:
    """Forecast init times as `filter_time_range` picks them (reference dataloader/utils.py:517-597): for every month
    present in `times`, `num_samples_per_month` days evenly spaced from the 1st up to (excluding) the last day of the
    month, at 00 and 12 UTC, never beyond the last available time.  `times`: datetimes (or anything
    datetime.fromisoformat / numpy datetime64 convertible); returns a sorted list of datetime objects."""
    import calendar
    from datetime import datetime

    import numpy as np

    def as_dt(t):
        if isinstance(t, datetime):
            return t
        if isinstance(t, np.datetime64):
            return t.astype("datetime64[s]").astype(datetime)
        return datetime.fromisoformat(str(t))

    ts = [as_dt(t) for t in times]
    if enforce_year is not None:
        ts = [t for t in ts if t.year == int(enforce_year)]
    if not ts:
        return []
    last = max(as_dt(t) for t in times)
    picked = []
    for yr, mo in sorted({(t.year, t.month) for t in ts}):
        days = np.round(np.linspace(1, calendar.monthrange(yr, mo)[1], num_samples_per_month, endpoint=False)).astype(int)
        days[0] = 1
        for day in days:
            for hour in (0, 12):
                dt = datetime(yr, mo, int(day), hour)
                if dt <= last:
                    picked.append(dt)
    return sorted(picked)
