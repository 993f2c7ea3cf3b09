"""Edgelist parser (reference `_io.py:132-295`)."""


def read_from_edgelist(df, **kwargs):
    raise NotImplementedError("vimure_b200.io.read_from_edgelist")
