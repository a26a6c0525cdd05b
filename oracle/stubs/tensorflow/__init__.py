"""Import-only placeholder; the Keras dataset download (tools.py:25-33) is never called."""


class _NoData:
    def __getattr__(self, name):
        raise RuntimeError("tensorflow is not available: datasets cannot be downloaded here")


class keras:  # noqa: N801
    datasets = _NoData()
