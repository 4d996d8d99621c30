from .upstream import Featurizer, S3PRLUpstream, UpstreamDownstreamModel  # noqa: F401
