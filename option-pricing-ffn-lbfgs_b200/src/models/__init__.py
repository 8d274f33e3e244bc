"""Pricing models (B200 COS pricer)."""
