"""TEST INFRASTRUCTURE — attribute-only stand-in for `xarray` so that the unmodified reference modules
(`ladcast/pipelines/utils.py:10`, `ladcast/evaluate/utils.py:6`, `ladcast/dataloader/utils.py:8`) import.
None of the hot-path functions touch xarray objects; anything that does raises here."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise NotImplementedError("xarray is not installed; the oracle shim only provides the names")


class Dataset(_Unavailable):
    pass


class DataArray(_Unavailable):
    pass


def open_zarr(*a, **k):
    raise NotImplementedError("xarray is not installed")


def merge(*a, **k):
    raise NotImplementedError("xarray is not installed")
