"""Empty stand-in: the reference's deprecated plotting imports are never exercised."""
