"""Minimal stand-in for the `easydict` package (absent from this image).

Only used by tests/golden/make_golden.py so that the reference's cfg.py can be
imported unmodified from /root/reference when the golden fixtures are minted.
"""


class EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
