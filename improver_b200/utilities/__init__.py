"""Host-side helpers mirroring improver/utilities for the neighbourhood path."""
