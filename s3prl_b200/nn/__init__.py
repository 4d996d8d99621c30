from .upstream import S3PRLUpstream  # noqa: F401
