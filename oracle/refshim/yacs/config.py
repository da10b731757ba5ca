"""Minimal CfgNode: attribute-access dict + merge_from_file, enough for the reference config."""
import copy
import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        self.__dict__["_new_allowed"] = new_allowed
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v, new_allowed=new_allowed) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def _merge(self, other):
        for k, v in other.items():
            if k in self and isinstance(self[k], CfgNode) and isinstance(v, dict):
                self[k]._merge(v)
            elif k in self or self._new_allowed:
                self[k] = CfgNode(v, new_allowed=True) if isinstance(v, dict) else v
            else:
                raise KeyError("Non-existent config key: {}".format(k))

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        for full_key, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        pass

    def defrost(self):
        pass
