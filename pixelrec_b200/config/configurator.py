"""YAML -> Config, same semantics as the reference's REC/config/configurator.py:13-180:
later files override earlier ones (:73-79), `1e-12` parses as float (:32-46), missing keys read as None
(:148-152), derived keys MODEL_INPUT_TYPE / eval_type / valid_metric_bigger (:98-120), topk validation
(:122-133).  The shipped reference yaml files (IDNet/sasrec.yaml, overall/ID.yaml, ...) load unchanged.
"""
import re

import yaml

from ..evaluator.metrics import metric_types, smaller_metrics
from ..utils.utils import get_model

_FLOAT_RE = re.compile(
    r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
        |\.[0-9_]+(?:[eE][-+][0-9]+)?
        |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\.[0-9_]*
        |[-+]?\.(?:inf|Inf|INF)
        |\.(?:nan|NaN|NAN))$""", re.X)


def _yaml_loader():
    class Loader(yaml.FullLoader):
        pass
    Loader.add_implicit_resolver("tag:yaml.org,2002:float", _FLOAT_RE, list("-+0123456789."))
    return Loader


class Config:
    def __init__(self, config_file_list=None, config_dict=None):
        self.final_config_dict = {}
        loader = _yaml_loader()
        for path in (config_file_list or []):
            with open(path, "r", encoding="utf-8") as f:
                d = yaml.load(f.read(), Loader=loader)
            if d:
                self.final_config_dict.update(d)
        if config_dict:
            self.final_config_dict.update(config_dict)
        if "model" not in self.final_config_dict:
            raise ValueError("config must name a `model`")
        self.model_class = get_model(self.final_config_dict["model"])
        self._derive()

    def _derive(self):
        d = self.final_config_dict
        if hasattr(self.model_class, "input_type"):
            d["MODEL_INPUT_TYPE"] = self.model_class.input_type
        metrics = d.get("metrics", ["Recall", "NDCG"])
        if isinstance(metrics, str):
            metrics = [metrics]
        d["metrics"] = metrics
        kinds = set()
        for m in metrics:
            if m.lower() not in metric_types:
                raise NotImplementedError(f"There is no metric named '{m}'")
            kinds.add(metric_types[m.lower()])
        if len(kinds) > 1:
            raise RuntimeError("Ranking metrics and value metrics can not be used at the same time.")
        d["eval_type"] = kinds.pop()
        vm = d.get("valid_metric", "NDCG@10")
        d["valid_metric"] = vm
        d["valid_metric_bigger"] = vm.split("@")[0].lower() not in smaller_metrics
        topk = d.get("topk", [5, 10])
        if isinstance(topk, int):
            topk = [topk]
        if not isinstance(topk, list):
            raise TypeError(f"The topk [{topk}] must be a integer, list")
        for k in topk:
            if k <= 0:
                raise ValueError(f"topk must be a positive integer or a list of positive integers, but get `{k}`")
        d["topk"] = topk

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        self.final_config_dict[key] = value

    def __getitem__(self, item):
        return self.final_config_dict.get(item, None)

    def __getattr__(self, item):
        d = self.__dict__.get("final_config_dict")
        if d is None:
            raise AttributeError("'Config' object has no attribute 'final_config_dict'")
        if item in d:
            return d[item]
        raise AttributeError(f"'Config' object has no attribute '{item}'")

    def __contains__(self, key):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        return key in self.final_config_dict

    def __str__(self):
        return "\n".join(f"{k} = {v}" for k, v in self.final_config_dict.items())

    __repr__ = __str__
