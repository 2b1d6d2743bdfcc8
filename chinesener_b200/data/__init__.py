"""Host-side feature pipeline: tokenizers and the feature-dict builder that feed the engine."""
