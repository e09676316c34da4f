# mirrors lib/hough_voting_layer/__init__.py of the reference
