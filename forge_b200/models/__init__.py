"""Host-side mirror of the reference's ``models`` package for the render / rotate / fusion path.

Module, class, method and ``state_dict`` key names follow the reference (models/*.py) so its
scripts can import this package in place of theirs; the arithmetic on volumes runs in
libforge_b200.so.
"""
