"""Empty stand-in so the reference's module-level `import matplotlib...` succeeds
(test infrastructure only; base.py:4, sls_base.py:4, utils.py:6-8 import it at module level)."""


def __getattr__(name):
    def _dummy(*a, **k):
        return None
    return _dummy
