"""Empty stand-in for cufflinks (plotting; never exercised)."""
