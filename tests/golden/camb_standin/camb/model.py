"""Placeholder for `camb.model` (the reference only touches it in get_cmb_cls, off the hot path)."""
NonLinear_both = "NonLinear_both"
NonLinear_none = "NonLinear_none"
