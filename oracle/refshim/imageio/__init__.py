"""Stub for imageio (imported, unused on the hot path)."""
