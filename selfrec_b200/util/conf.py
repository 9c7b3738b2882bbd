"""ModelConf with the contract of util/conf.py:5-28: a flat YAML mapping, `conf[key]` exits with the
reference's message when the key is missing, an unreadable file raises IOError."""
import os

import yaml


def _load_yaml(path):
    if not os.path.exists(path):
        print("Config file is not found!")
        raise IOError
    try:
        with open(path, "r") as fh:
            return yaml.safe_load(fh)
    except yaml.YAMLError as exc:
        print(f"Error in configuration file: {exc}")
        raise IOError


class ModelConf(object):
    def __init__(self, file=None, config=None):
        self.config = dict(config) if config is not None else _load_yaml(file)

    def read_configuration(self, file):
        self.config = _load_yaml(file)

    def contain(self, key):
        return key in self.config

    def __getitem__(self, item):
        try:
            return self.config[item]
        except KeyError:
            print("Parameter " + item + " is not found in the configuration file!")
            exit(-1)
