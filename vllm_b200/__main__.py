from vllm_b200.entrypoints.cli import main

main()
