from .homo_benchmark import HomoBenchmark  # noqa: F401
