import functools
import inspect


def register_to_config(init):
    """diffusers.configuration_utils.register_to_config: record ctor kwargs on self.config (after __init__)."""
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        init(self, *args, **kwargs)
        sig = inspect.signature(init)
        bound = sig.bind_partial(self, *args, **kwargs)
        for name, p in sig.parameters.items():
            if name == "self" or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
                continue
            val = bound.arguments.get(name, p.default)
            setattr(self.config, name, val)
    return inner
