from pathlib import Path


def posix_path(*parts) -> str:
    """Join path components and return a POSIX-style string (GPErks/serialization/path.py)."""
    return Path(*map(str, parts)).as_posix()
