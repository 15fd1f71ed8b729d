"""Empty stand-in for matplotlib."""
