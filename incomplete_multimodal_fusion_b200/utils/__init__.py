"""Callers either side of the hot path: the device half of the reference's dataset transforms."""
