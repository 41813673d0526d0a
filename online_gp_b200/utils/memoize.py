"""Per-object memoisation with explicit invalidation — the protocol of ``gpytorch.utils.memoize``
(``cached(name=...)`` / ``pop_from_cache`` / ``CachingError``) that ``FixedNoiseOnlineSKIGP`` relies on
(``online_gp/models/batched_fixed_noise_online_gp.py:335,344,351,358,364,369,406-414``)."""
import functools


class CachingError(RuntimeError):
    pass


def cached(method=None, name=None):
    if method is None:
        return functools.partial(cached, name=name)
    cache_name = name if name is not None else method.__name__

    @functools.wraps(method)
    def g(self, *args, **kwargs):
        store = self.__dict__.setdefault("_memoize_cache", {})
        if cache_name not in store:
            store[cache_name] = method(self, *args, **kwargs)
        return store[cache_name]

    return g


def is_in_cache(obj, name):
    return name in obj.__dict__.get("_memoize_cache", {})


def pop_from_cache(obj, name):
    try:
        return obj.__dict__["_memoize_cache"].pop(name)
    except KeyError:
        raise CachingError(f"Object does not have item {name} stored in cache.")


def clear_cache(obj):
    obj.__dict__["_memoize_cache"] = {}
