"""Config objects of the reference (misc/utils.py:13-61): JSON -> attribute bag with a ``.dict`` view that the
operator functions both read and mutate (defaults, float() coercions, ``global_step``)."""
import json


class Params():
    """Loads hyper-parameters from a nnet_conf/*.json file (misc/utils.py:13-41); the files load unchanged."""

    def __init__(self, json_path):
        self.update(json_path)

    def save(self, json_path):
        with open(json_path, 'w') as f:
            json.dump(self.__dict__, f, indent=4)

    def update(self, json_path):
        with open(json_path) as f:
            params = json.load(f)
            self.__dict__.update(params)

    @property
    def dict(self):
        return self.__dict__


class ParamsPlain():
    """Manual parameter bag (misc/utils.py:44-61)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def dict(self):
        return self.__dict__


def substring_in_list(s, varlist):
    """misc/utils.py:315-330."""
    if varlist is None:
        return False
    for v in varlist:
        if v in s:
            return True
    return False
