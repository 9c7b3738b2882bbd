"""Minimal mirror of data/loader.py (FileIO): `user item weight` triples per line."""
import os
from collections.abc import Sequence


def _parse(file):
    data = []
    with open(file) as f:
        for line in f:
            parts = line.strip().split(" ")
            data.append([parts[0], parts[1], float(parts[2])])
    return data


class TripleFile(Sequence):
    """What FileIO.load_data_set returns: behaves like the reference's list of [user, item, weight] (parsed on
    first use), and remembers its path so that Interaction can hand the file to the native builder
    (csrc/dataset.cpp) instead of looping over a million Python lists."""

    def __init__(self, path):
        self.path = os.fspath(path)
        self._rows = None

    def _load(self):
        if self._rows is None:
            self._rows = _parse(self.path)
        return self._rows

    def __len__(self):
        return len(self._load())

    def __getitem__(self, k):
        return self._load()[k]

    def __iter__(self):
        return iter(self._load())

    def __eq__(self, other):
        return list(self) == list(other)


class FileIO(object):
    """data/loader.py's FileIO.  The graph format goes through TripleFile (and from there to the native builder);
    the other readers are plain restatements so that aliasing this module never takes a format away."""

    @staticmethod
    def write_file(dir, file, content, op="w"):
        os.makedirs(dir, exist_ok=True)
        with open(dir + file, op) as f:  # plain concatenation, as the reference does (loader.py:14)
            f.writelines(content)

    @staticmethod
    def delete_file(file_path):
        if os.path.exists(file_path):
            os.remove(file_path)

    @staticmethod
    def load_data_set(file, rec_type="graph"):
        if rec_type == "graph":
            return TripleFile(file)
        if rec_type == "sequential":  # `seq_id:item item ...` per line (loader.py:34-40); not on the hot path
            with open(file) as f:
                return {head: tail.split() for head, tail in (line.strip().split(":")[:2] for line in f)}
        raise ValueError(f"unknown dataset type {rec_type!r}")

    @staticmethod
    def load_user_list(file):
        print("loading user List...")
        with open(file) as f:
            return [line.strip().split()[0] for line in f]

    @staticmethod
    def load_social_data(file):
        print("loading social data...")
        out = []
        with open(file) as f:
            for line in f:
                parts = line.strip().split(" ")
                out.append([parts[0], parts[1], 1 if len(parts) < 3 else float(parts[2])])
        return out
