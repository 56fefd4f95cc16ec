"""Empty stand-in: the unconditional path never touches open_clip."""
