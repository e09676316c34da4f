# mirrors lib/average_distance_loss/__init__.py of the reference
