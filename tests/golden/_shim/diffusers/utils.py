from dataclasses import fields


class BaseOutput:
    """Minimal diffusers.utils.BaseOutput: attribute + key access on a dataclass."""
    def __getitem__(self, k):
        return getattr(self, k)

    def keys(self):
        return [f.name for f in fields(self)]
