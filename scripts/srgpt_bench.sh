#!/bin/bash
# SpatialRGPT-Bench over the sm_100a path: one process per GPU, each answering its chunk of the annotation file, answers
# merged at the end (the launch pattern of the reference's scripts/srgpt/eval/srgpt_bench.sh:9-49).
#   bash scripts/srgpt_bench.sh <model_path> <annotation.json> <image_folder> [conv_mode] [out_dir]
set -euo pipefail
MODEL_PATH=$1
ANNOTATIONS=$2
IMAGE_FOLDER=$3
CONV_MODE=${4:-llama_3}
OUT_DIR=${5:-eval_output/$(basename "$MODEL_PATH")/SpatialRGPT-Bench}

IFS=',' read -ra GPUS <<< "${CUDA_VISIBLE_DEVICES:-0}"
CHUNKS=${#GPUS[@]}
mkdir -p "$OUT_DIR"
ROOT=$(cd "$(dirname "$0")/.." && pwd)

for IDX in $(seq 0 $((CHUNKS - 1))); do
  CUDA_VISIBLE_DEVICES=${GPUS[$IDX]} PYTHONPATH="$ROOT:${PYTHONPATH:-}" python -m llava.eval.eval_spatial \
    --model-path "$MODEL_PATH" --annotation-file "$ANNOTATIONS" --image-folder "$IMAGE_FOLDER" \
    --answers-file "$OUT_DIR/${CHUNKS}_${IDX}.jsonl" --num-chunks "$CHUNKS" --chunk-idx "$IDX" \
    --temperature 0 --conv-mode "$CONV_MODE" &
done
wait

: > "$OUT_DIR/merge.jsonl"
for IDX in $(seq 0 $((CHUNKS - 1))); do
  cat "$OUT_DIR/${CHUNKS}_${IDX}.jsonl" >> "$OUT_DIR/merge.jsonl"
done
echo "answers: $OUT_DIR/merge.jsonl ($(wc -l < "$OUT_DIR/merge.jsonl") records)"
