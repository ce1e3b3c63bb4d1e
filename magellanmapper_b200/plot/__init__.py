"""Mirror of the reference sub-package of the same name, hot path only."""
