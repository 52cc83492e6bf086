"""Resolution of the reference's Hydra callable strings (``tester.sampler_callable``, ``network.callable``,
``diff_params.callable``) the way ``dnnlib.call_func_by_name`` does it (utils/dnnlib/util.py:292-297, used at
utils/setup.py:49,55 and testing/blind_bwe_tester.py:214): import the longest importable module prefix, walk the
remaining attributes, call with the caller's keyword arguments."""
import importlib


def get_obj_by_name(name):
    parts = name.split(".")
    for i in range(len(parts) - 1, 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ImportError:
            continue
        for attr in parts[i:]:
            obj = getattr(obj, attr)
        return obj
    raise ImportError(name)


def call_func_by_name(*args, func_name=None, **kwargs):
    assert func_name is not None
    fn = get_obj_by_name(func_name)
    assert callable(fn)
    return fn(*args, **kwargs)
