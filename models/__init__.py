"""`models` package at the repo root so the reference's own driver line
`from models.graphcnn import *` (main.py:9) picks up the B200-native classes unchanged."""
