"""Feature / output label files: one label per line (GPErks/serialization/labels.py:1-3)."""
from pathlib import Path


def read_labels_from_file(labels_file_path):
    return Path(labels_file_path).read_text().splitlines()
