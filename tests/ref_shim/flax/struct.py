from _dataclass import dataclass          # noqa: F401
