from ._core import SharedVar as SharedVariable, Function, function  # noqa: F401
