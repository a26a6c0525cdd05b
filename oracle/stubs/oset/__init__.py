"""Placeholder for the ``oset`` package (ordered set used only for tags)."""


class oset(list):
    pass
