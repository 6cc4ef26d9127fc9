"""YAML -> nested attribute object (same semantics as the reference's config/load_config.py:6-43)."""

import os

import yaml

CONFIG_DIR = os.path.dirname(os.path.abspath(__file__))


class Config:
    @staticmethod
    def from_yaml(filepath):
        with open(filepath, "r") as fp:
            return Config.from_dict(yaml.safe_load(fp))

    @staticmethod
    def from_dict(data):
        cfg = Config()
        for name, value in data.items():
            setattr(cfg, name, Config.from_dict(value) if isinstance(value, dict) else value)
        return cfg

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Config) else v) for k, v in self.__dict__.items()}

    def get(self, name, default=None):
        return getattr(self, name, default)
