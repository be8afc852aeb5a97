import json

import numpy


class NumpyEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, numpy.integer):
            return int(obj)
        if isinstance(obj, numpy.floating):
            return float(obj)
        if isinstance(obj, numpy.ndarray):
            return obj.tolist()
        return super().default(obj)


def save_json(dct, filename):
    with open(filename, "w") as f:
        json.dump(dct, f, cls=NumpyEncoder, indent=4)


def load_json(filename):
    """Lists come back as numpy arrays (GPErks/utils/jsonfiles.py:27-40)."""
    def hook(d):
        return {k: (numpy.array(v) if isinstance(v, list) else v) for k, v in d.items()}
    with open(filename, "r") as f:
        return json.load(f, object_hook=hook)
