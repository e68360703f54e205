#!/bin/bash
tools/gpu_round.sh r2e
tools/sanitize.sh racecheck
