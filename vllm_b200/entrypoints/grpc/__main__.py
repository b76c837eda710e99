from vllm_b200.entrypoints.grpc import main

main()
