"""Minimal stand-in for the `omegaconf` package (TEST INFRASTRUCTURE ONLY).

The reference's hot-path modules import `omegaconf` only for
`OmegaConf.merge/create/to_container/set_struct/set_readonly`
(/root/reference/gluefactory/models/matchers/lightglue.py:341,
/root/reference/gluefactory/models/utils/losses.py:37,
/root/reference/gluefactory/models/base_model.py:65-86).  omegaconf is not
installed in this image, so `oracle/make_golden.py` puts this directory on
`sys.path` to import the UNMODIFIED reference and generate golden vectors.
Nothing in the product path imports this.
"""
import copy


class DictConfig(dict):
    """Attribute-access dict with recursive wrapping."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = _wrap(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, DictConfig):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
        return ListConfig(_wrap(x) for x in v)
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    return v


def _merge_into(a, b):
    for k, v in b.items():
        if k in a and isinstance(a[k], dict) and isinstance(v, dict):
            _merge_into(a[k], v)
        else:
            a[k] = _wrap(copy.deepcopy(v))
    return a


class OmegaConf:
    @staticmethod
    def create(d=None):
        return _wrap(copy.deepcopy(d if d is not None else {}))

    @staticmethod
    def merge(*confs):
        out = DictConfig()
        for c in confs:
            if c is None:
                continue
            _merge_into(out, _wrap(c))
        return out

    @staticmethod
    def to_container(c, resolve=True):
        return _unwrap(c)

    @staticmethod
    def set_struct(c, flag):
        return None

    @staticmethod
    def set_readonly(c, flag):
        return None

    @staticmethod
    def is_dict(c):
        return isinstance(c, dict)

    @staticmethod
    def load(path):
        import yaml

        with open(path) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def from_cli(args=None):
        return DictConfig()

    @staticmethod
    def save(c, path):
        import yaml

        with open(path, "w") as f:
            yaml.safe_dump(_unwrap(c), f)
