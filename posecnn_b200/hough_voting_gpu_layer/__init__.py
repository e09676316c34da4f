# mirrors lib/hough_voting_gpu_layer/__init__.py of the reference
