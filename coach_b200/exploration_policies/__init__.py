"""Exploration policies of the device acting path."""
