"""Host-side mirror of the reference's ``STMiner/Algorithm`` package for the hot path."""
