from socks_b200.sampling.transform import transpose_sample  # noqa: F401
