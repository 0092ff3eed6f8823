"""Shim: the reference imports `timm.models.layers` only for `trunc_normal_` and `DropPath`
(/root/reference/models/_layers.py:6).  timm is not installed in this image."""
