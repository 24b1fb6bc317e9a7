"""Training-target synthesis (confidence maps, part affinity fields)."""
